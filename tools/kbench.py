#!/usr/bin/env python
"""Kernel-only timing of one configuration with the library K3D_LIB points at (tuning aid).

    K3D_LIB=/path/libk3d_variant.so python tools/kbench.py c3 [reps] [scale]

Prints search/solve device times (best of reps) and a checksum of the results so that
variants can be compared with each other."""
import os
import sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
import k3d_configs
from pygeostatistics_b200 import Krige3d, _native

name = sys.argv[1] if len(sys.argv) > 1 else "c3"
reps = int(sys.argv[2]) if len(sys.argv) > 2 else 3
scale = float(sys.argv[3]) if len(sys.argv) > 3 else 1.0
par, data, sec = k3d_configs.config(name, scale)
k = Krige3d(par, data=data, secondary=sec)
k.verbose = False
k.prepare()
P = k._params_struct()
plane = par['nx'] * par['ny']
nodes = plane * par['nz']
host = k._host_inputs()
ext = k._external_grid()
d = k._upload(host, None if ext is None else ext.pin_memory())
out = k._alloc_outputs(nodes, neighbours=False)
best = [1e9, 1e9]
for r in range(reps + 1):
    k._launch(P, d, 0, par['nz'], out)
    torch.cuda.synchronize()
    li = _native.last_launch()
    if r:
        best = [min(best[0], li['search_ms']), min(best[1], li['solve_ms'])]
est = out['est']
ok = ~torch.isnan(est)
print("%s lib=%s search %.2f ms solve %.2f ms total %.2f ms | shape %dx%d smem %d | est sum %.12e var sum %.12e nan %d" % (
    name, os.path.basename(_native.LIB_PATH), best[0], best[1], best[0] + best[1], li['grid'], li['block'],
    li['smem_bytes'], float(est[ok].sum()), float(out['var'][ok].sum()), int((~ok).sum())))
