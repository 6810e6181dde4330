import os, sys, time
sys.path.insert(0, os.getcwd())
import numpy as np, torch, torch.distributed as dist
import k3d_configs
from pygeostatistics_b200 import Krige3d
import pygeostatistics_b200.krige3d as K
nccl = bool(os.environ.get("DIAG_NCCL"))
local = int(os.environ.get("LOCAL_RANK", "0")) if nccl else 0
torch.cuda.set_device(local)
if nccl:
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
else:
    dist.init_process_group("gloo")
rank = dist.get_rank()
if rank == 0:
    os.system("df -h /dev/shm | tail -1")
par, data, sec = k3d_configs.config("c3", 1.0)
for rep in range(3):
    k = Krige3d(par, data=data); k.verbose = False
    t0 = time.time(); k.kt3d(); dt = time.time() - t0
    sg = list(K._SHARED.values())
    print(rank, rep, "dt %.1f ms" % (dt * 1e3), {a: round(b, 4) for a, b in k.timings.items()},
          "shared:", [(s.ok, s.nbytes, len(s.registered)) for s in sg], flush=True)
    del k
dist.destroy_process_group()
