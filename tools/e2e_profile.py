#!/usr/bin/env python
"""Host-side profile of Krige3d(...).kt3d() (tuning aid): cProfile of the second call."""
import cProfile
import os
import pstats
import sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import k3d_configs
from pygeostatistics_b200 import Krige3d

name = sys.argv[1] if len(sys.argv) > 1 else "c3"
par, data, sec = k3d_configs.config(name, 1.0)


def once():
    k = Krige3d(par, data=data, secondary=sec)
    k.verbose = False
    k.kt3d()
    return k


once()
pr = cProfile.Profile()
pr.enable()
k = once()
pr.disable()
print(k.timings)
pstats.Stats(pr).sort_stats("cumulative").print_stats(35)
