"""The krige2d goldens (tests/golden/krige2d_reference.npz) as test cases."""
import json
import os

import numpy as np


def krige2d_cases(golden_dir):
    """(tag, Krige2d parameter dict, data columns, golden est, golden var) of the runs of the
    UNMODIFIED reference krige2d.py kept in krige2d_reference.npz."""
    g = np.load(os.path.join(golden_dir, "krige2d_reference.npz"))
    c1 = json.loads(str(g['c1_par'][0]))
    x, y, v = g['c1_xyv']
    for tag, isk in (("sk", 0), ("ok", 1)):
        yield "c1_" + tag, dict(c1, isk=isk), {'x': x, 'y': y, 'por': v}, \
            g['c1_est_' + tag], g['c1_var_' + tag]
    x, y, v = g['syn_xyv']
    for tag in ("ok_nested", "sk_nested", "ok_ndmin3"):
        yield tag, json.loads(str(g['syn_%s_par' % tag][0])), {'x': x, 'y': y, 'v': v}, \
            g['syn_%s_est' % tag], g['syn_%s_var' % tag]
    for i in range(int(g['nrand'][0])):
        x, y, v = g['rand%d_xyv' % i]
        yield "rand%d" % i, json.loads(str(g['rand%d_par' % i][0])), {'x': x, 'y': y, 'v': v}, \
            g['rand%d_est' % i], g['rand%d_var' % i]
