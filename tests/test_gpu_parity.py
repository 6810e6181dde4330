"""GPU parity: the CUDA path (through the C ABI, via Krige3d) against the CPU oracle and the
committed golden vectors.  Tolerances (BASELINE.json north_star):
  neighbours  identical ordered lists of sample ids (bit exact)
  estimate    1e-9 relative (+ 1e-12 * max|data| absolute: estimates can cancel to ~0)
  variance    1e-9 relative or 1e-12 absolute (variances are ~0 at data-coincident nodes)
  NaN pattern identical
"""
import json
import os

import numpy as np
import pytest

import k3d_configs
from oracle import kt3d_oracle as orc

torch = pytest.importorskip("torch")
pytestmark = pytest.mark.gpu

EST_RTOL, VAR_RTOL, VAR_ATOL = 1e-9, 1e-9, 1e-12


def gpu_run(par, data, sec, neighbours=True):
    from pygeostatistics_b200 import Krige3d
    k = Krige3d(par, data=data, secondary=sec)
    k.verbose = False
    k.kt3d(neighbours=neighbours)
    return k


def oracle_run(par, data, sec, node_list=None, threads=8):
    if data is None:
        names, cols = orc.read_geoeas(par['datafl'])
    else:
        names, cols = list(data.keys()), dict(data)
        if 'z' not in cols:
            cols['z'] = np.zeros_like(cols['x'])
    prop = [n for n in names if n not in ('x', 'y', 'z')]
    st = orc.prepare(par, cols, prop)
    return st, orc.run(st, extgrid=sec, node_list=node_list, threads=threads)


def compare(k, st, ores, node_list=None, vmax=1.0, check_neighbours=True):
    est, var, nbr, cnt = ores
    sel = slice(None) if node_list is None else node_list
    gest, gvar = k.estimation[sel], k.estimation_variance[sel]
    assert np.array_equal(st['sort_index'], k.searcher.sort_index)
    if check_neighbours:
        assert np.array_equal(k.neighbour_count[sel], cnt)
        assert np.array_equal(k.neighbours_sorted_space[sel], nbr.astype(np.int64))
    assert np.array_equal(np.isnan(gest), np.isnan(est))
    assert np.array_equal(np.isnan(gvar), np.isnan(var))
    ok = ~np.isnan(est)
    if ok.any():
        derr = np.abs(gest[ok] - est[ok])
        assert np.all(derr <= EST_RTOL * np.abs(est[ok]) + 1e-12 * vmax), \
            "estimate: max abs err %g" % derr.max()
        verr = np.abs(gvar[ok] - var[ok])
        assert np.all((verr <= VAR_RTOL * np.abs(var[ok])) | (verr <= VAR_ATOL)), \
            "variance: max abs err %g" % verr.max()
    return int(ok.sum())


# ---------------------------------------------------------------- config 1 (shipped)
@pytest.mark.parametrize("name,tag", [("c1", "sk"), ("c1ok", "ok")])
def test_config1_against_unmodified_reference_golden(golden_dir, name, tag):
    par, data, sec = k3d_configs.config(name)
    k = gpu_run(par, data, sec)
    g = np.load(os.path.join(golden_dir, "c1_reference.npz"))
    e, v = g['est_' + tag], g['var_' + tag]
    assert not np.isnan(k.estimation).any()
    assert np.max(np.abs(k.estimation - e) / np.abs(e)) < EST_RTOL
    assert np.max(np.abs(k.estimation_variance - v)) < 1e-11
    # neighbour SETS of the unmodified reference (it returns them index-ordered, D4)
    assert np.array_equal(k.neighbour_count, g['nclose'])
    sets = np.sort(np.where(k.neighbours >= 0, k.neighbours, 1 << 30), axis=1)
    sets = np.where(sets == 1 << 30, -1, sets)
    width = g['nbr_sets'].shape[1]
    full = np.full((sets.shape[0], width), -1, np.int64)
    full[:, :min(width, sets.shape[1])] = sets[:, :width]
    assert np.array_equal(full, g['nbr_sets'].astype(np.int64))
    # and the ordered lists against the oracle
    st, ores = oracle_run(par, data, sec)
    compare(k, st, ores, vmax=20.0)
    # super-block state the reference exposes
    assert k.searcher.nsbtosr == g['tmpl'].shape[1]
    assert np.array_equal(np.stack([k.searcher.ixsbtosr, k.searcher.iysbtosr,
                                    k.searcher.izsbtosr]), g['tmpl'])
    assert np.array_equal(k.searcher.nisb, g['nisb'])
    assert np.array_equal(k.rotmat, g['rotmat'])


@pytest.mark.parametrize("name", ["small_sk_sph", "small_ok_nested_aniso",
                                  "small_ok_gauss_hole_2d", "small_sk_power",
                                  "small_ok_ndmin"])
def test_small_goldens_from_unmodified_reference(golden_dir, name):
    g = np.load(os.path.join(golden_dir, name + ".npz"))
    par = json.loads(str(g['par'][0]))
    data = {'x': g['x'], 'y': g['y'], 'z': g['z'], 'v': g['v']}
    k = gpu_run(par, data, None)
    assert np.array_equal(np.isnan(k.estimation), np.isnan(g['est']))
    ok = ~np.isnan(g['est'])
    scale = max(1.0, np.max(np.abs(g['est'][ok])))
    assert np.max(np.abs(k.estimation[ok] - g['est'][ok])) < 1e-9 * scale
    assert np.max(np.abs(k.estimation_variance[ok] - g['var'][ok])) < \
        1e-9 * max(1.0, np.max(np.abs(g['var'][ok])))
    assert np.array_equal(k.neighbour_count, g['nclose'])


# ---------------------------------------------------------------- scaled synthetic configs
@pytest.mark.parametrize("name,scale", [("c2", 0.3), ("c3", 0.18), ("c4", 0.18), ("c5", 0.09)])
def test_scaled_configs_against_oracle(name, scale):
    par, data, sec = k3d_configs.config(name, scale)
    k = gpu_run(par, data, sec)
    st, ores = oracle_run(par, data, sec)
    n = compare(k, st, ores, vmax=float(np.max(np.abs(data['v']))))
    assert n > 0.5 * len(k.estimation)


def test_unordered_selection_path_and_plane_runs(monkeypatch):
    """Production path: without the neighbour lists the kept set is not sorted (rows of the
    system come in selection order) and the lists cross through pooled scratch; with a tiny
    scratch cap the slab is processed in several runs of z-planes.  Estimates must still
    match the oracle."""
    par, data, sec = k3d_configs.config("c3", 0.2)
    st, ores = oracle_run(par, data, sec)
    k = gpu_run(par, data, sec, neighbours=False)
    compare(k, st, ores, vmax=float(np.max(np.abs(data['v']))), check_neighbours=False)
    monkeypatch.setenv("K3D_SCRATCH_CAP_MB", "1")      # ~2 planes per run at this size
    from pygeostatistics_b200 import _native
    n0 = _native.last_launch()['launches']
    k2 = gpu_run(par, data, sec, neighbours=False)
    assert _native.last_launch()['launches'] - n0 > 4   # several search+solve pairs
    compare(k2, st, ores, vmax=float(np.max(np.abs(data['v']))), check_neighbours=False)
    assert np.allclose(k2.estimation, k.estimation, rtol=1e-11, atol=1e-12, equal_nan=True)


# ---------------------------------------------------------------- full size, sampled nodes
@pytest.mark.parametrize("name", ["c2", "c3", "c4"])
def test_full_size_sampled_nodes_against_oracle(name):
    par, data, sec = k3d_configs.config(name, 1.0)
    k = gpu_run(par, data, sec)
    nodes = par['nx'] * par['ny'] * par['nz']
    rng = np.random.default_rng(7)
    node_list = np.sort(rng.choice(nodes, 6000, replace=False))
    # always include the grid corners and a face centre (clipped neighbourhoods)
    node_list[:3] = [0, par['nx'] - 1, nodes - 1]
    node_list = np.unique(node_list)
    st, ores = oracle_run(par, data, sec, node_list=node_list)
    compare(k, st, ores, node_list=node_list, vmax=float(np.max(np.abs(data['v']))))
    assert np.all(k.status[~np.isnan(k.estimation)] == 0)


def test_full_size_linearity_c3():
    """Size-independent property at the full 256^3: the OK estimate is linear in the data
    and the variance does not depend on them: est(a*v + b) = a*est(v) + b."""
    par, data, sec = k3d_configs.config("c3", 1.0)
    k1 = gpu_run(par, data, sec, neighbours=False)
    d2 = dict(data)
    d2['v'] = 2.5 * data['v'] + 7.0
    k2 = gpu_run(par, d2, sec, neighbours=False)
    ok = ~np.isnan(k1.estimation)
    assert ok.mean() > 0.99
    assert np.allclose(k2.estimation[ok], 2.5 * k1.estimation[ok] + 7.0, rtol=1e-10, atol=1e-10)
    assert np.array_equal(k1.estimation_variance[ok], k2.estimation_variance[ok])


# ---------------------------------------------------------------- edge cases
def _edge_par(**kw):
    par = dict(k3d_configs.BASE)
    par.update(nx=12, ny=10, nz=6, radius_hmax=3.0, radius_hmin=3.0, radius_vert=3.0,
               aa_hmax=[6.0], aa_hmin=[6.0], aa_vert=[6.0], ndmax=8)
    par.update(kw)
    return par


def _edge_data(n, seed=11, ext=(12, 10, 6)):
    rng = np.random.default_rng(seed)
    p = rng.uniform(0, 1, (n, 3)) * np.array(ext)
    return {'x': p[:, 0].copy(), 'y': p[:, 1].copy(), 'z': p[:, 2].copy(),
            'v': rng.normal(size=n)}


@pytest.mark.parametrize("kw,n", [
    (dict(ikrige=0), 1),                       # a single sample: one-sample closed form
    (dict(ikrige=1), 1),                       # OK with na=1 <= mdt: everything unestimated
    (dict(ikrige=0, ndmin=3), 25),             # ndmin rejections
    (dict(ikrige=1, noct=1, ndmax=8), 200),    # octant search, one per octant
    (dict(ikrige=1, noct=2, ndmax=5), 200),    # octant search truncated by ndmax
    (dict(ikrige=1, ndmax=64), 300),           # largest ndmax of the build
    (dict(ikrige=0, radius_hmax=9.0, radius_hmin=9.0, radius_vert=9.0, ndmax=12), 2500),  # list prune
    (dict(ikrige=1, idrift=[True, True, True] + [False] * 6, ndmax=12), 300),  # linear drift
    (dict(ikrige=1, itrend=True, idrift=[True] + [False] * 8, ndmax=12), 300),  # trend
    (dict(ikrige=0, nxdis=2, nydis=2, nzdis=2, ndmax=10), 200),                # block SK
    (dict(ikrige=1, nxdis=3, nydis=1, nzdis=1, ndmax=10), 200),                # 3 block points
    (dict(ikrige=1, nxdis=1, nydis=2, nzdis=1, ndmax=40), 300),                # 2 points, 2 groups
    (dict(ikrige=1, nxdis=5, nydis=5, nzdis=4, ndmax=12), 200),                # 100 points
    (dict(ikrige=0, it=[3], c0=0.3, cc=[0.7]), 150),                           # gaussian
    (dict(ikrige=1, it=[4], c0=0.1, cc=[0.4], aa_hmax=[1.2], aa_hmin=[1.2], aa_vert=[1.2]), 150),
    (dict(ikrige=1, it=[5], c0=0.3, cc=[0.7]), 150),                           # hole effect
    (dict(ikrige=0, nz=1, zmn=0.0, radius_vert=1.0), 60),                      # 2-D grid
])
def test_edge_cases_against_oracle(kw, n):
    par = _edge_par(**kw)
    ext = (par['nx'], par['ny'], par['nz'] if par['nz'] > 1 else 0)
    data = _edge_data(n, ext=ext)
    k = gpu_run(par, data, None)
    st, ores = oracle_run(par, data, None)
    compare(k, st, ores, vmax=float(np.max(np.abs(data['v']))))


def test_dense_tile_falls_back_to_global_walk():
    """More candidates in a tile than the shared-memory stage holds."""
    par = _edge_par(ikrige=1, nx=8, ny=8, nz=4, radius_hmax=6.0, radius_hmin=6.0,
                    radius_vert=6.0, ndmax=16)
    data = _edge_data(9000, seed=5, ext=(8, 8, 4))
    k = gpu_run(par, data, None)
    st, ores = oracle_run(par, data, None)
    compare(k, st, ores, vmax=float(np.max(np.abs(data['v']))))


def test_ked_and_nonstationary_sk():
    par, data, sec = k3d_configs.config("c5", 0.06)
    for ikrige in (3, 2):
        p = dict(par, ikrige=ikrige, nxdis=1, nydis=1, nzdis=1)
        k = gpu_run(p, data, sec)
        st, ores = oracle_run(p, data, sec)
        compare(k, st, ores, vmax=float(np.max(np.abs(data['v']))))
    # external values outside [tmin, tmax) are trimmed (krige3d.py:339)
    p = dict(par, ikrige=3, tmax=float(np.median(sec)))
    k = gpu_run(p, data, sec)
    st, ores = oracle_run(p, data, sec)
    compare(k, st, ores, vmax=float(np.max(np.abs(data['v']))))
    assert np.isnan(k.estimation).sum() >= (sec >= p['tmax']).sum()


def test_exact_interpolation_and_far_field():
    """Known answers: a node coinciding with a sample returns the datum with ~0 variance;
    SK beyond the search radius is unestimated; OK weights reproduce a constant field."""
    par = _edge_par(ikrige=1, ndmax=12)
    data = _edge_data(150)
    data['x'][0], data['y'][0], data['z'][0] = 3.5, 4.5, 2.5  # node (3,4,2)
    k = gpu_run(par, data, None)
    node = (2 * par['ny'] + 4) * par['nx'] + 3
    assert abs(k.estimation[node] - data['v'][0]) < 1e-9
    assert abs(k.estimation_variance[node]) < 1e-9
    const = dict(data)
    const['v'] = np.full_like(data['v'], 3.25)
    k2 = gpu_run(par, const, None)
    ok = ~np.isnan(k2.estimation)
    assert np.allclose(k2.estimation[ok], 3.25, rtol=1e-10)


def test_file_based_drop_in(tmp_path):
    """Krige3d(param_file) with on-disk par/gslib/secfl files and the gslib writer."""
    from pygeostatistics_b200 import Krige3d, SpatialData
    par, data, sec = k3d_configs.config("c5", 0.05)
    pfile = k3d_configs.write_files(str(tmp_path), par, data, sec)
    k = Krige3d(pfile)
    k.read_data()
    k.verbose = False
    k.kriging()
    # identical to an in-memory run on the values the reader parsed ...
    parsed = SpatialData(k.datafl).vr
    from pygeostatistics_b200.gslib_io import read_grid_column
    k_mem = gpu_run(par, {n: parsed[n] for n in ('x', 'y', 'z', 'v', 'sec')},
                    read_grid_column(k.extfl, k.iextve), neighbours=False)
    assert np.array_equal(k.estimation, k_mem.estimation, equal_nan=True)
    # ... and to the original columns up to the text round trip (the reference's
    # pd.read_csv call uses pandas' fast float parser, which is not correctly rounded)
    k_org = gpu_run(par, data, sec)
    assert np.allclose(k.estimation, k_org.estimation, rtol=1e-9, atol=1e-12, equal_nan=True)
    out = str(tmp_path / "out.dat")
    k.write_output(out, unest=-999.0)
    lines = open(out).read().split("\n")
    assert lines[1] == "2" and lines[2] == "Estimate" and lines[3] == "EstimationVariance"
    vals = np.loadtxt(out, skiprows=4)
    assert vals.shape == (par['nx'] * par['ny'] * par['nz'], 2)
    ok = ~np.isnan(k.estimation)
    assert np.array_equal(vals[ok, 0], k.estimation[ok])
    assert np.all(vals[~ok] == -999.0)


def test_host_buffer_c_abi_entry():
    """k3d_krige_slab_host: the same path with host pointers (what a ctypes binding from
    the reference would call), slab by slab."""
    import ctypes as C
    from pygeostatistics_b200 import Krige3d, _native
    par, data, sec = k3d_configs.config("c3", 0.12)
    k = gpu_run(par, data, sec)
    k.brick_size = (1, 1, 1)    # the hand-made inputs below carry the one-block template
    P = k._params_struct()
    s = k.searcher
    arrs = dict(sx=np.ascontiguousarray(k.vr['x']), sy=np.ascontiguousarray(k.vr['y']),
                sz=np.ascontiguousarray(k.vr['z']), sv=np.ascontiguousarray(k.vr['v']),
                sbstart=s.sbstart, runs=np.ascontiguousarray(s.runs.reshape(-1)),
                nodex0=s.nodex0, nodey0=s.nodey0, nodez0=s.nodez0,
                dbxyz=np.concatenate([k.xdb, k.ydb, k.zdb]))
    I = _native.K3dInputs()
    I.nd = len(arrs['sx'])
    for name in ('sx', 'sy', 'sz', 'sv', 'sbstart', 'runs', 'nodex0', 'nodey0', 'nodez0', 'dbxyz'):
        setattr(I, name, arrs[name].ctypes.data)
    I.nruns = len(s.runs)
    I.ntmpl = s.nsbtosr
    plane = par['nx'] * par['ny']
    est = np.empty(plane * par['nz']); var = np.empty_like(est)
    half = par['nz'] // 2
    for a, b in ((0, half), (half, par['nz'])):
        O = _native.K3dOutputs()
        O.est = est[a * plane:].ctypes.data
        O.var = var[a * plane:].ctypes.data
        _native.check(_native.lib().k3d_krige_slab_host(C.byref(P), C.byref(I), a, b, C.byref(O)))
    # not bit-identical: each slab walks its own node path, so the rows of a system can
    # be ordered differently (covariance reuse); the values agree to rounding
    assert np.allclose(est, k.estimation, rtol=1e-11, atol=1e-12, equal_nan=True)
    assert np.allclose(var, k.estimation_variance, rtol=1e-11, atol=1e-12, equal_nan=True)
    # k3d_krige_points_host: cross-validation of the same data through host pointers
    kx = Krige3d(dict(par, option=1), data=data)
    kx.verbose = False
    kx.kt3d()
    x, y, z = (np.ascontiguousarray(data[c]) for c in 'xyz')
    blk = s.block_of_locations(x, y, z)
    order = np.argsort(blk, kind='stable')
    nsup = s.nxsup * s.nysup * s.nzsup
    pa = dict(x=x[order], y=y[order], z=z[order],
              start=np.concatenate([[0], np.cumsum(np.bincount(blk, minlength=nsup))]).astype(np.int32))
    Q = _native.K3dPoints()
    Q.npts = len(x)
    for name in ('x', 'y', 'z', 'start'):
        setattr(Q, name, pa[name].ctypes.data)
    Q.exclude_coincident = 1
    pe = np.empty(len(x)); pv = np.empty(len(x))
    O = _native.K3dOutputs()
    O.est, O.var = pe.ctypes.data, pv.ctypes.data
    _native.check(_native.lib().k3d_krige_points_host(C.byref(P), C.byref(I), C.byref(Q), C.byref(O)))
    back = np.empty(len(x)); back[order] = pe
    assert np.array_equal(back, kx.estimation, equal_nan=True)
    back[order] = pv
    assert np.array_equal(back, kx.estimation_variance, equal_nan=True)


# ---------------------------------------------------------------- option 1 / 2: listed locations
def _oracle_points(par, data, px, py, pz, ext=None, koption=1):
    names, cols = list(data.keys()), dict(data)
    prop = [n for n in names if n not in ('x', 'y', 'z')]
    st = orc.prepare(par, cols, prop)
    return st, orc.run_points(st, px, py, pz, ext=ext, koption=koption, threads=8)


def _compare_points(k, st, ores, vmax):
    est, var, nbr, cnt = ores
    assert np.array_equal(k.neighbour_count, cnt)
    assert np.array_equal(k.neighbours_sorted_space, nbr.astype(np.int64))
    assert np.array_equal(np.isnan(k.estimation), np.isnan(est))
    ok = ~np.isnan(est)
    derr = np.abs(k.estimation[ok] - est[ok])
    assert np.all(derr <= EST_RTOL * np.abs(est[ok]) + 1e-12 * vmax), derr.max()
    verr = np.abs(k.estimation_variance[ok] - var[ok])
    assert np.all((verr <= VAR_RTOL * np.abs(var[ok])) | (verr <= VAR_ATOL)), verr.max()
    return int(ok.sum())


@pytest.mark.parametrize("name,scale", [("c3", 0.4), ("c2", 0.5), ("c4", 0.3)])
def test_cross_validation_against_oracle(name, scale):
    """option 1 (krige3d.py:200-208, 310-332, 359-363): every datum re-estimated from the
    others.  c3 exercises the octant rule (the datum itself takes its place in its octant and
    is dropped afterwards), c2 the plain ndmax rule, c4 the drift rows."""
    from pygeostatistics_b200 import Krige3d
    par, data, sec = k3d_configs.config(name, scale)
    par = dict(par, option=1)
    k = Krige3d(par, data=data)
    k.verbose = False
    k.kt3d(neighbours=True)
    st, ores = _oracle_points(par, data, data['x'], data['y'], data['z'], koption=1)
    n = _compare_points(k, st, ores, vmax=float(np.max(np.abs(data['v']))))
    assert n > 0.9 * len(data['x'])
    # a datum never appears among its own neighbours, and is not reproduced exactly
    own = np.arange(len(data['x']))[:, None]
    assert not (k.neighbours == own).any()
    assert np.nanmax(np.abs(k.estimation - data['v'])) > 1e-3
    assert np.array_equal(k.true_values, data['v'])
    # the production path (no neighbour lists, unordered selection) gives the same numbers
    k2 = Krige3d(par, data=data)
    k2.verbose = False
    k2.kt3d()
    assert np.allclose(k2.estimation, k.estimation, rtol=1e-10, atol=1e-12, equal_nan=True)


def test_jackknife_locations_and_grid_equivalence(tmp_path):
    """option 2: estimation at the records of `jackfl`.  Locations on grid nodes must give
    what the grid run gives there; arbitrary locations (some outside the grid: clamped to the
    edge super block) must match the oracle; the file round trip works."""
    from pygeostatistics_b200 import Krige3d
    par, data, sec = k3d_configs.config("c3", 0.2)
    nx, ny, nz = par['nx'], par['ny'], par['nz']
    rng = np.random.default_rng(5)
    idx = np.sort(rng.choice(nx * ny * nz, 3000, replace=False))
    iz = idx // (nx * ny); iy = (idx - iz * nx * ny) // nx; ix = idx - iz * nx * ny - iy * nx
    gx, gy, gz = par['xmn'] + ix * par['xsiz'], par['ymn'] + iy * par['ysiz'], par['zmn'] + iz * par['zsiz']
    kg = gpu_run(par, data, sec)
    jpar = dict(par, option=2)
    kj = Krige3d(jpar, data=data, jackknife={'x': gx, 'y': gy, 'z': gz, 'v': np.zeros(len(gx))})
    kj.verbose = False
    kj.kt3d(neighbours=True)
    assert np.array_equal(kj.neighbours, kg.neighbours[idx])
    assert np.allclose(kj.estimation, kg.estimation[idx], rtol=1e-10, atol=1e-12, equal_nan=True)
    assert np.allclose(kj.estimation_variance, kg.estimation_variance[idx], rtol=1e-9, atol=1e-12,
                       equal_nan=True)
    # arbitrary locations, a few outside the grid extent, a few on data
    m = 4000
    px = rng.uniform(-2.0, nx + 2.0, m); py = rng.uniform(-2.0, ny + 2.0, m); pz = rng.uniform(0.0, nz, m)
    px[:50], py[:50], pz[:50] = data['x'][:50], data['y'][:50], data['z'][:50]
    jk = {'x': px, 'y': py, 'z': pz, 'v': rng.normal(size=m)}
    # through files: jackfl read by column name, outfl written per location
    from pygeostatistics_b200.gslib_io import SpatialData
    jf = tmp_path / "jack.dat"
    with open(jf, "w") as f:
        f.write("jack\n4\nx\ny\nz\nv\n")
        for r in zip(px, py, pz, jk['v']):
            f.write("\t".join(repr(float(c)) for c in r) + "\n")
    fpar = dict(jpar, jackfl=str(jf), outfl=str(tmp_path / "xv.out"))
    kf = Krige3d(fpar, data=data)
    kf.verbose = False
    kf.kt3d(neighbours=True)
    # (the oracle gets the coordinates as the reader parsed them: pandas' default float
    # parser, the one the reference uses, may differ from repr() in the last bit)
    fx, fy, fz = kf.locations.T
    st, ores = _oracle_points(jpar, data, fx, fy, fz, koption=2)
    assert _compare_points(kf, st, ores, vmax=float(np.max(np.abs(data['v'])))) > 0.9 * m
    # in memory: the 50 locations that ARE data lose that datum, octant slot included
    km = Krige3d(jpar, data=data, jackknife=jk)
    km.verbose = False
    km.kt3d(neighbours=True)
    st, ores = _oracle_points(jpar, data, px, py, pz, koption=2)
    assert _compare_points(km, st, ores, vmax=float(np.max(np.abs(data['v'])))) > 0.9 * m
    assert not (km.neighbours[:50] == np.arange(50)[:, None]).any()
    kf.write_output()
    out = SpatialData(str(tmp_path / "xv.out")).vr
    assert out.dtype.names[:3] == ('X', 'Y', 'Z') and len(out) == m
    # (the reader's float parser is pandas' fast one, exact to a few ulp)
    assert np.allclose(out['Estimate'], kf.estimation, rtol=1e-12, atol=1e-14, equal_nan=True)
    assert np.allclose(out['Error: est-true'], kf.estimation - jk['v'], rtol=1e-9, atol=1e-12,
                       equal_nan=True)


def test_cross_validation_ked_external_drift():
    """option 1 with kriging with an external drift: the external value at each location
    comes from the data's secondary column (krige3d.py:329-330)."""
    from pygeostatistics_b200 import Krige3d
    par, data, sec = k3d_configs.config("c5", 0.06)
    par = dict(par, option=1, nxdis=1, nydis=1, nzdis=1)
    k = Krige3d(par, data=data)
    k.verbose = False
    k.kt3d(neighbours=True)
    names = [n for n in data if n not in ('x', 'y', 'z')]
    st, ores = _oracle_points(par, data, data['x'], data['y'], data['z'], ext=data[names[1]], koption=1)
    assert _compare_points(k, st, ores, vmax=float(np.max(np.abs(data[names[0]])))) > 0.5 * len(data['x'])


# ---------------------------------------------------------------- krige2d front end
def _write_geoeas(path, cols):
    names = list(cols.keys())
    with open(path, "w") as f:
        f.write("data\n%d\n%s\n" % (len(names), "\n".join(names)))
        for row in zip(*[cols[n] for n in names]):
            f.write("  ".join(repr(float(c)) for c in row) + "\n")


def test_krige2d_front_end_against_unmodified_reference(golden_dir, tmp_path):
    """Krige2d(par).read_data()/kd2d() -- the reference's krige2d.py interface on the krige3d
    kernels -- against runs of the unmodified reference (tests/golden/krige2d_reference.npz)."""
    from pygeostatistics_b200.krige2d import Krige2d
    from k2d_cases import krige2d_cases as _krige2d_cases
    for tag, par, cols, est, var in _krige2d_cases(golden_dir):
        data = tmp_path / (tag + ".gslib")
        _write_geoeas(data, cols)
        pf = tmp_path / (tag + ".par")
        json.dump(dict(par, datafl=str(data)), open(pf, "w"))
        k = Krige2d(str(pf))
        k.verbose = False
        k.read_data()
        k.kd2d()
        assert k.estimation.shape == (par['nx'], par['ny'])
        assert np.array_equal(np.isnan(k.estimation), np.isnan(est)), tag
        ok = ~np.isnan(est)
        rel = np.abs(k.estimation[ok] - est[ok]) / np.maximum(np.abs(est[ok]), 1e-3)
        assert rel.max() < EST_RTOL, (tag, rel.max())
        assert np.max(np.abs(k.estimation_variance[ok] - var[ok])) < 1e-10, tag


@pytest.mark.parametrize("upd", [dict(), dict(isk=0, nxdis=3, nydis=2), dict(it=[3, 1], ndmax=12)])
def test_krige2d_dense_cases_against_oracle(upd, tmp_path):
    """More than ndmax samples in range (the reference's neighbour rule is order dependent
    there, defect K1: the ndmax nearest are kept), block kriging (crashes in the reference,
    K5) and a Gaussian structure (K2): the repaired restatement is the oracle."""
    from pygeostatistics_b200.krige2d import Krige2d
    rng = np.random.default_rng(9)
    n = 1500
    cols = {'x': rng.uniform(0, 60, n), 'y': rng.uniform(0, 45, n), 'v': rng.normal(size=n)}
    par = dict(datafl="", icolx=0, icoly=1, icolvr=2, tmin=-1e21, tmax=1e21, idbg=0,
               dbgfl="kb2d.dbg", outfl="out.dat", nx=60, xmn=0.5, xsiz=1.0, ny=45, ymn=0.5,
               ysiz=1.0, nxdis=1, nydis=1, ndmin=1, ndmax=16, radius=6.0, isk=1, skmean=0.0,
               nst=2, c0=0.1, it=[2, 1], cc=[0.4, 0.5], azm=[30.0, 110.0], a_max=[12.0, 20.0],
               a_min=[5.0, 9.0])
    par.update(upd)
    data = tmp_path / "d.gslib"
    _write_geoeas(data, cols)
    par['datafl'] = str(data)
    k = Krige2d(par)
    k.verbose = False
    k.kd2d()
    c3 = {'x': k.vr['x'], 'y': k.vr['y'], 'z': k.vr['z'], 'v': k.vr['v']}
    st = orc.prepare(k.krige3d_params(), c3, ['v'])
    e, v, _, cnt = orc.run(st, threads=8)
    assert cnt.max() == par['ndmax']
    e = e.reshape(par['ny'], par['nx']).T
    v = v.reshape(par['ny'], par['nx']).T
    assert np.array_equal(np.isnan(k.estimation), np.isnan(e))
    ok = ~np.isnan(e)
    assert np.all(np.abs(k.estimation[ok] - e[ok]) <= EST_RTOL * np.abs(e[ok]) + 1e-11)
    assert np.all(np.abs(k.estimation_variance[ok] - v[ok]) <= VAR_RTOL * np.abs(v[ok]) + 1e-11)
