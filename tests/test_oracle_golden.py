"""Pins the CPU oracle (oracle/) against vectors produced by the UNMODIFIED reference
(tests/golden/make_golden.py).  No GPU, no /root/reference needed."""
import ctypes as C
import json
import os

import numpy as np
import pytest

from oracle import kt3d_oracle as orc


def _cols(x, y, z, v):
    return ['x', 'y', 'z', 'v'], dict(x=x, y=y, z=z, v=v)


def _orig_sets(st, nbr_idx, nbr_cnt, width):
    out = np.full((len(nbr_cnt), width), -1, np.int64)
    for k, n in enumerate(nbr_cnt):
        out[k, :n] = np.sort(st['sort_index'][nbr_idx[k, :n]])
    return out


@pytest.fixture(scope="module")
def c1(golden_dir):
    g = np.load(os.path.join(golden_dir, "c1_reference.npz"))
    par = orc.read_par(os.path.join(golden_dir, "c1_test_krige3d.par"))
    names, cols = orc.read_geoeas(os.path.join(golden_dir, "c1_test.gslib"))
    return g, par, names, cols


@pytest.mark.parametrize("ikrige,tag", [(0, "sk"), (1, "ok")])
def test_config1_matches_unmodified_reference(c1, ikrige, tag):
    g, par, names, cols = c1
    par = dict(par, ikrige=ikrige)
    st = orc.prepare(par, cols, [n for n in names if n not in 'xyz'])
    est, var, nbr, cnt = orc.run(st)
    # super-block state
    assert np.array_equal(st['nisb'], g['nisb'])
    assert np.array_equal(np.stack([st['tx'], st['ty'], st['tz']]), g['tmpl'])
    assert np.array_equal(st['rotmat'], g['rotmat'])
    assert st['maxcov'] == g['scalars'][0] and st['resc'] == g['scalars'][1]
    # 1 sample per super block on this data set, so the permutation is unique
    assert np.array_equal(st['sort_index'], g['sort_index'])
    # neighbour sets: the unmodified search returns them index-ordered (D4) but
    # complete (< ndmax in range), so compare as sets of original ids
    assert np.array_equal(cnt, g['nclose'])
    assert np.array_equal(_orig_sets(st, nbr, cnt, 32), g['nbr_sets'].astype(np.int64))
    e, v = g['est_' + tag], g['var_' + tag]
    assert not np.isnan(est).any()
    assert np.max(np.abs(est - e) / np.abs(e)) < 1e-12
    assert np.max(np.abs(var - v)) < 1e-12


def test_rotation_rescale_maxcov(golden_dir):
    g = np.load(os.path.join(golden_dir, "kernels_reference.npz"))
    par = json.loads(str(g['rot_par'][0]))
    x = np.array([1.0, 2.0]); cols = dict(x=x, y=x.copy(), z=x.copy(), v=x.copy())
    st = orc.prepare(par, cols, ['v'])
    assert np.array_equal(st['rotmat'], g['rot_rotmat'])          # bit exact
    assert st['maxcov'] == g['rot_scalars'][0]
    assert st['resc'] == g['rot_scalars'][1]
    assert st['radsqd'] == g['rot_scalars'][2]


def test_cova3_left_right(golden_dir):
    g = np.load(os.path.join(golden_dir, "kernels_reference.npz"))
    par = json.loads(str(g['rot_par'][0]))
    x = np.array([1.0, 2.0]); cols = dict(x=x, y=x.copy(), z=x.copy(), v=x.copy())
    st = orc.prepare(par, cols, ['v'])
    P, L = st['P'], orc.lib()
    cov = np.array([L.orc_cova3(C.byref(P), *a, *b) for a, b in zip(g['cova_p1'], g['cova_p2'])])
    # libm vs Numba's exp/cos/sqrt: <= a few ulp
    assert np.max(np.abs(cov - g['cova_out'])) < 5e-16 * np.max(np.abs(g['cova_out'])) * 8
    assert np.all(cov[:10] == P.maxcov)
    # power model set: structures (power 1.5, spherical) using rotmat 0 and 1
    P2 = orc.OrcParams.from_buffer_copy(P)
    P2.nst = 2
    P2.it[0], P2.it[1] = 4, 1
    P2.cc[0], P2.cc[1] = 0.5, 0.4
    P2.aa[0], P2.aa[1] = 1.5, 6.0
    P2.maxcov = 0.1 + 999 + 0.4
    covp = np.array([L.orc_cova3(C.byref(P2), *a, *b) for a, b in zip(g['cova_p1'], g['cova_p2'])])
    assert np.max(np.abs(covp - g['cova_pow_out']) / np.abs(g['cova_pow_out'])) < 1e-14
    # left_side / right_side
    D = C.POINTER(C.c_double)
    xa, ya, za = (np.ascontiguousarray(a) for a in g['ls_xyz'])
    na = len(xa)
    for neq, key in ((na, 'ls_sk'), (na + 1, 'ls_ok')):
        left = np.empty((neq, neq))
        L.orc_left_side(C.byref(P), na, neq, xa.ctypes.data_as(D), ya.ctypes.data_as(D),
                        za.ctypes.data_as(D), left.ctypes.data_as(D))
        assert np.allclose(left, g[key], rtol=1e-14, atol=1e-15, equal_nan=True)
    z1 = np.zeros(1)
    right = np.empty(na + 1)
    L.orc_right_side(C.byref(P), na, na + 1, xa.ctypes.data_as(D), ya.ctypes.data_as(D),
                     za.ctypes.data_as(D), z1.ctypes.data_as(D), z1.ctypes.data_as(D),
                     z1.ctypes.data_as(D), right.ctypes.data_as(D))
    assert np.allclose(right, g['rs_point'], rtol=1e-14, atol=0)
    xa, ya, za = (np.ascontiguousarray(a) for a in g['rs_block_xyz'])
    xdb, ydb, zdb = (np.ascontiguousarray(a) for a in g['rs_block_db'])
    P3 = orc.OrcParams.from_buffer_copy(P)
    P3.ndb = len(xdb)
    L.orc_right_side(C.byref(P3), na, na + 1, xa.ctypes.data_as(D), ya.ctypes.data_as(D),
                     za.ctypes.data_as(D), xdb.ctypes.data_as(D), ydb.ctypes.data_as(D),
                     zdb.ctypes.data_as(D), right.ctypes.data_as(D))
    assert np.allclose(right, g['rs_block'], rtol=1e-14, atol=0)


def test_pickup_template(golden_dir):
    g = np.load(os.path.join(golden_dir, "kernels_reference.npz"))
    nxs, nys, nzs, sx, sy, sz, radsqd = g['pick_args']
    rot = np.ascontiguousarray(g['rot_rotmat'][-1].ravel())
    L = orc.lib()
    L.orc_pickup.restype = C.c_int
    n = 4096
    tx = np.zeros(n, np.int32); ty = tx.copy(); tz = tx.copy()
    I = C.POINTER(C.c_int32)
    nt = L.orc_pickup(int(nxs), int(nys), int(nzs), C.c_double(sx), C.c_double(sy),
                      C.c_double(sz), rot.ctypes.data_as(C.POINTER(C.c_double)),
                      C.c_double(radsqd), tx.ctypes.data_as(I), ty.ctypes.data_as(I),
                      tz.ctypes.data_as(I))
    assert nt == g['pick_tmpl'].shape[1]
    assert np.array_equal(np.stack([tx[:nt], ty[:nt], tz[:nt]]), g['pick_tmpl'])


def test_setup_binning_and_getindx(golden_dir):
    g = np.load(os.path.join(golden_dir, "kernels_reference.npz"))
    geo = json.loads(str(g['setup_geo'][0]))
    par = dict(json.loads(str(g['rot_par'][0])), **geo)
    x, y, z = g['setup_xyz']
    st = orc.prepare(par, dict(x=x, y=y, z=z, v=x.copy()), ['v'])
    # the fixture forced MAXSB=(18,11,5) == what krige3d.py:167-181 derives for 37x23x11
    assert st['nsup'] == tuple(int(v) for v in g['setup_sup'][:3])
    assert st['sizsup'] == tuple(g['setup_sup'][3:6])
    assert st['mnsup'] == tuple(g['setup_sup'][6:9])
    assert np.array_equal(st['nisb'], g['setup_nisb'])
    assert np.array_equal(st['block_of_sample'], g['setup_block_of_sample'])
    assert np.array_equal(st['lutx'], g['getindx_x'])
    assert np.array_equal(st['luty'], g['getindx_y'])
    assert np.array_equal(st['lutz'], g['getindx_z'])


@pytest.mark.parametrize("name", ["small_sk_sph", "small_ok_nested_aniso",
                                  "small_ok_gauss_hole_2d", "small_sk_power",
                                  "small_ok_ndmin"])
def test_small_full_runs_match_unmodified_reference(golden_dir, name):
    g = np.load(os.path.join(golden_dir, name + ".npz"))
    par = json.loads(str(g['par'][0]))
    names, cols = _cols(g['x'], g['y'], g['z'], g['v'])
    st = orc.prepare(par, cols, ['v'])
    est, var, nbr, cnt = orc.run(st, threads=2)
    assert np.array_equal(cnt, g['nclose'])
    assert np.array_equal(_orig_sets(st, nbr, cnt, par['ndmax']), g['nbr_sets'].astype(np.int64))
    assert np.array_equal(np.isnan(est), np.isnan(g['est']))
    assert np.array_equal(np.isnan(var), np.isnan(g['var']))
    ok = ~np.isnan(est)
    scale = max(1.0, np.max(np.abs(g['est'][ok])))
    assert np.max(np.abs(est[ok] - g['est'][ok])) < 1e-10 * scale
    assert np.max(np.abs(var[ok] - g['var'][ok])) < 1e-10 * max(1.0, np.max(np.abs(g['var'][ok])))


def test_oracle_locations_mode_properties():
    """Cross-validation / jackknife restatement (krige3d.py:310-332, 359-363; the reference
    crashes there, D11 -- parity unpinned by the reference): at grid nodes it IS the grid
    loop; at the data it interpolates exactly unless the datum is excluded."""
    import k3d_configs
    par, data, sec = k3d_configs.config("c3", 0.2)
    st = orc.prepare(par, dict(data), ['v'])
    nx, ny, nz = par['nx'], par['ny'], par['nz']
    idx = np.arange(0, nx * ny * nz, 11)
    iz = idx // (nx * ny); iy = (idx - iz * nx * ny) // nx; ix = idx - iz * nx * ny - iy * nx
    px = par['xmn'] + ix * par['xsiz']; py = par['ymn'] + iy * par['ysiz']; pz = par['zmn'] + iz * par['zsiz']
    a = orc.run(st, node_list=idx, threads=4)
    for kopt in (0, 2):  # no datum sits on a node of this grid: exclusion changes nothing
        b = orc.run_points(st, px, py, pz, koption=kopt, threads=4)
        for u, v in zip(a, b):
            assert np.array_equal(u, v, equal_nan=True)
    x, y, z, v = data['x'], data['y'], data['z'], data['v']
    e0, v0, n0, c0 = orc.run_points(st, x, y, z, koption=0, threads=4)
    assert np.max(np.abs(e0 - v)) < 1e-9 and np.max(np.abs(v0)) < 1e-9
    e1, v1, n1, c1 = orc.run_points(st, x, y, z, koption=1, threads=4)
    srt = np.argsort(st['sort_index'])           # original row -> sorted position
    assert not (n1 == srt[:, None]).any()        # never its own neighbour
    assert np.nanmin(v1) > 1e-3 and np.nanmax(np.abs(e1 - v)) > 1e-2
    assert np.all(c1 <= par['ndmax'])


from k2d_cases import krige2d_cases as _krige2d_cases  # noqa: E402


def test_oracle_drives_krige2d_against_unmodified_reference(golden_dir):
    """The krige2d front end maps onto the krige3d path (pygeostatistics_b200/krige2d.py):
    the oracle, driven with that mapping, reproduces the unmodified reference krige2d.py
    (nested anisotropic variograms, SK / OK, ndmin, nodes with a single neighbour: kb2d's
    ordinary-kriging rule) wherever its order-dependent neighbour rule cannot fire."""
    from pygeostatistics_b200.krige2d import Krige2d
    for tag, par, cols, est, var in _krige2d_cases(golden_dir):
        k = Krige2d(par)
        p3 = k.krige3d_params()
        cols = dict(cols, z=np.zeros_like(cols['x']))
        names = [n for n in cols if n not in ('x', 'y', 'z')]
        st = orc.prepare(p3, cols, names)
        e, v, _, _ = orc.run(st, threads=4)
        e = e.reshape(par['ny'], par['nx']).T
        v = v.reshape(par['ny'], par['nx']).T
        assert np.array_equal(np.isnan(e), np.isnan(est)), tag
        ok = ~np.isnan(est)
        assert np.max(np.abs(e[ok] - est[ok]) / np.maximum(np.abs(est[ok]), 1e-3)) < 1e-9, tag
        assert np.max(np.abs(v[ok] - var[ok])) < 1e-10, tag
