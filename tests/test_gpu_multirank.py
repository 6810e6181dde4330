"""The multi-rank branch of Krige3d.kt3d() on hardware: two ranks (gloo; both on cuda:0, which
NCCL would refuse) krige their z-slabs, copy them into the host buffer they share, and every
rank must see the whole grid equal to the oracle's."""
import os
import subprocess
import sys

import pytest

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))

_WORKER = r'''
import os, sys
sys.path.insert(0, %r)
import numpy as np, torch, torch.distributed as dist
import k3d_configs
from oracle import kt3d_oracle as orc
from pygeostatistics_b200 import Krige3d
torch.cuda.set_device(0)
dist.init_process_group("gloo")
rank, world = dist.get_rank(), dist.get_world_size()
for name, scale in (("c3", 0.16), ("c5", 0.07)):
    par, data, sec = k3d_configs.config(name, scale)
    st = orc.prepare(par, dict(data), [n for n in data if n not in ("x", "y", "z")])
    oe, ov, onbr, ocnt = orc.run(st, extgrid=sec, threads=4)
    for rep in range(2):                       # the second run re-uses the shared buffer
        k = Krige3d(par, data=data, secondary=sec)
        k.verbose = False
        k.kt3d(neighbours=(rep == 1))
        assert k.estimation.shape == oe.shape
        assert np.array_equal(np.isnan(k.estimation), np.isnan(oe)), (name, rank)
        ok = ~np.isnan(oe)
        vmax = float(np.max(np.abs(data["v"])))
        assert np.all(np.abs(k.estimation[ok] - oe[ok]) <= 1e-9 * np.abs(oe[ok]) + 1e-12 * vmax)
        dv = np.abs(k.estimation_variance[ok] - ov[ok])
        assert np.all((dv <= 1e-9 * np.abs(ov[ok])) | (dv <= 1e-12))
        assert np.all(k.status[ok] == 0)
        if rep == 1:
            assert np.array_equal(k.neighbours_sorted_space, onbr.astype(np.int64))
            assert np.array_equal(k.neighbour_count, ocnt)
        del k
if rank == 0:
    print("MULTIRANK_OK")
dist.destroy_process_group()
'''


@pytest.mark.parametrize("nproc", [2, 3])
def test_kt3d_under_a_process_group(tmp_path, nproc):
    script = tmp_path / "worker.py"
    script.write_text(_WORKER % ROOT)
    port = 33500 + os.getpid() % 2000 + nproc
    out = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1",
                          "--nproc-per-node=%d" % nproc, "--master-addr", "127.0.0.1",
                          "--master-port", str(port), str(script)], capture_output=True,
                         text=True, timeout=600)
    assert out.returncode == 0, out.stdout[-1500:] + out.stderr[-3000:]
    assert "MULTIRANK_OK" in out.stdout
