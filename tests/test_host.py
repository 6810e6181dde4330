"""CPU-only tests of the host side: the C-ABI library loads and exports every symbol of
include/k3d.h, parameter / file handling mirrors the reference, the super-block index
matches the reference's golden state, and the z-slab sharding + gather works under a
world_size-2 gloo group."""
import ctypes as C
import json
import os
import re
import subprocess
import sys

import numpy as np
import pytest

import k3d_configs
from pygeostatistics_b200 import (Krige3d, SpatialData, SuperBlockSearcher, slab_ranges,
                                  write_gslib_grid, _native)
from pygeostatistics_b200.super_block import (template_runs, runs_from_offsets, bins_left,
                                                _edges_sample)

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_library_exports_every_declared_symbol():
    hdr = open(os.path.join(ROOT, "include", "k3d.h")).read()
    declared = set(re.findall(r"\b(k3d_[a-z0-9_]+)\s*\(", hdr))
    assert declared == set(_native.SYMBOLS)
    lib = _native.lib()
    for sym in declared:
        assert getattr(lib, sym) is not None
    assert b"sm_100a" in lib.k3d_version()
    # struct layouts agree with the header (sizes the C side was compiled with)
    assert C.sizeof(_native.K3dParams) % 8 == 0
    assert C.sizeof(_native.K3dOutputs) == 6 * 8
    # argument validation happens before any CUDA call: usable without a GPU
    P = _native.K3dParams(); I = _native.K3dInputs(); O = _native.K3dOutputs()
    assert lib.k3d_krige_slab(C.byref(P), C.byref(I), 0, 1, C.byref(O), None) == -1
    assert b"grid" in lib.k3d_last_error()
    P.nx = P.ny = P.nz = 4
    P.ndmax = 100
    assert lib.k3d_krige_slab(C.byref(P), C.byref(I), 0, 1, C.byref(O), None) == -1
    assert b"ndmax" in lib.k3d_last_error()


def test_struct_size_matches_compiled_header(tmp_path):
    src = tmp_path / "sz.c"
    src.write_text('#include <stdio.h>\n#include "k3d.h"\nint main(){printf("%zu %zu %zu %zu",'
                   'sizeof(K3dParams),sizeof(K3dInputs),sizeof(K3dOutputs),sizeof(K3dLaunchInfo));}')
    exe = tmp_path / "sz"
    subprocess.check_call(["gcc", "-I", os.path.join(ROOT, "include"), str(src), "-o", str(exe)])
    sizes = [int(v) for v in subprocess.check_output([str(exe)]).split()]
    assert sizes == [C.sizeof(_native.K3dParams), C.sizeof(_native.K3dInputs),
                     C.sizeof(_native.K3dOutputs), C.sizeof(_native.K3dLaunchInfo)]


def test_param_file_and_checks(tmp_path):
    par, _, _ = k3d_configs.config("c1")
    k = Krige3d(os.path.join(ROOT, "tests", "golden", "c1_test_krige3d.par"),
                data={'x': np.zeros(2), 'y': np.zeros(2), 'por': np.zeros(2)})
    assert (k.nx, k.ny, k.nz, k.ndmax, k.ktype) == (98, 79, 1, 30, 0)
    assert k.tmax == 1e21 and isinstance(k.tmax, float)
    assert k._2d and k.property_name == ['por'] and k.vr.dtype.names == ('x', 'y', 'por', 'z')
    for bad, exc in ((dict(radius_hmax=0), ValueError), (dict(nst=0), ValueError),
                     (dict(it=[7]), ValueError), (dict(it=[4], aa_hmax=[2.5]), ValueError),
                     (dict(idrift=[0] * 9), ValueError), (dict(ikrige=3, icolsec=0), ValueError),
                     (dict(ndmax=65), ValueError), (dict(option=3), ValueError),
                     (dict(aa_hmax=[0.0]), ValueError)):
        with pytest.raises(exc):
            Krige3d(dict(par, **bad), data={'x': np.zeros(2), 'y': np.zeros(2), 'v': np.zeros(2)})
    missing = dict(par)
    del missing['ndmax']
    with pytest.raises(KeyError):
        Krige3d(missing, data={'x': np.zeros(2), 'y': np.zeros(2), 'v': np.zeros(2)})
    with pytest.raises(FileNotFoundError):
        Krige3d(dict(par, datafl=str(tmp_path / "nope.gslib")))


def test_geoeas_reader_and_writer(tmp_path, golden_dir):
    d = SpatialData(os.path.join(golden_dir, "c1_test.gslib"))
    assert d.vr.shape == (85,) and d._2d and d.property_name == ['por']
    assert d.vr['x'][0] == 12100.0 and d.vr['por'][0] == 14.6515 and np.all(d.vr['z'] == 0)
    out = str(tmp_path / "o.dat")
    est = np.array([1.5, np.nan, 3.25]); var = np.array([0.1, np.nan, 0.0])
    write_gslib_grid(out, est, var, unest=-999.0)
    txt = open(out).read().split("\n")
    assert txt[1:4] == ["2", "Estimate", "EstimationVariance"]
    assert np.array_equal(np.loadtxt(out, skiprows=4), [[1.5, 0.1], [-999, -999], [3.25, 0.0]])


def test_host_setup_matches_reference_state(golden_dir):
    """rotmat / maxcov / resc / binning / node LUTs of the product's own host code against
    the unmodified reference's golden state (no GPU: the template is injected)."""
    g = np.load(os.path.join(golden_dir, "c1_reference.npz"))
    par, _, _ = k3d_configs.config("c1")
    k = Krige3d(par)
    k._preprocess(); k._set_rotation(); k._max_covariance(); k._rescaling()
    assert np.array_equal(k.rotmat, g['rotmat'])
    assert (k.maxcov, k.resc, k.radsqd) == tuple(g['scalars'][:3])
    s = SuperBlockSearcher()
    s.nx, s.xmn, s.xsiz, s.ny, s.ymn, s.ysiz = k.nx, k.xmn, k.xsiz, k.ny, k.ymn, k.ysiz
    s.nz, s.zmn, s.zsiz, s.vr, s.MAXSB = k.nz, k.zmn, k.zsiz, k.vr, k.const.MAXSB
    s.setup()
    assert (s.nxsup, s.nysup, s.nzsup) == tuple(int(v) for v in g['scalars'][3:6])
    assert (s.xsizsup, s.ysizsup, s.zsizsup) == tuple(g['scalars'][6:9])
    assert np.array_equal(s.nisb, g['nisb'])
    assert np.array_equal(s.sort_index, g['sort_index'])
    assert s.sbstart[0] == 0 and s.sbstart[-1] == 85
    # second fixture: binning on edges + Numba-flavoured node lookup
    kg = np.load(os.path.join(golden_dir, "kernels_reference.npz"))
    geo = json.loads(str(kg['setup_geo'][0]))
    s = SuperBlockSearcher()
    for key, v in geo.items():
        setattr(s, key, v)
    x, y, z = kg['setup_xyz']
    s.vr = np.rec.fromarrays([x, y, z], names='x,y,z')
    s.MAXSB = (18, 11, 5)
    s.setup()
    assert np.array_equal(s.nisb, kg['setup_nisb'])
    assert np.array_equal(s.block_of_sample, kg['setup_block_of_sample'])
    assert np.array_equal(s.lutx, kg['getindx_x'])
    assert np.array_equal(s.luty, kg['getindx_y'])
    assert np.array_equal(s.lutz, kg['getindx_z'])
    for lut, start in ((s.lutx, s.nodex0), (s.luty, s.nodey0), (s.lutz, s.nodez0)):
        for b in range(len(start) - 1):
            assert np.all(lut[start[b]:start[b + 1]] == b)
        assert start[-1] == len(lut)
    # template -> runs -> template round trip, reference ordering of the offsets
    tm = kg['pick_tmpl']
    nxs, nys, nzs = (int(v) for v in kg['pick_args'][:3])
    m = np.zeros((2 * nxs - 1, 2 * nys - 1, 2 * nzs - 1), bool)
    m[tm[0] + nxs - 1, tm[1] + nys - 1, tm[2] + nzs - 1] = True
    s.nxsup, s.nysup, s.nzsup = nxs, nys, nzs
    s.set_template_mask(m)
    assert np.array_equal(np.stack([s.ixsbtosr, s.iysbtosr, s.izsbtosr]), tm)
    runs = template_runs(m, nxs, nys, nzs)
    back = np.zeros_like(m)
    for dx0, dx1, dy, dz in runs:
        assert not back[dx0 + nxs - 1: dx1 + nxs, dy + nys - 1, dz + nzs - 1].any()
        back[dx0 + nxs - 1: dx1 + nxs, dy + nys - 1, dz + nzs - 1] = True
    assert np.array_equal(back, m)
    key = runs[:, 3].astype(int) * 10000 + runs[:, 2].astype(int) * 100 + runs[:, 0]
    assert np.all(np.diff(key) > 0)
    # the device route hands over offset lists instead of the mask: same runs
    s.set_template_offsets(tm[0], tm[1], tm[2])
    assert np.array_equal(s.runs, runs) and s.nsbtosr == tm.shape[1]
    assert np.array_equal(runs_from_offsets(tm[0][::-1], tm[1][::-1], tm[2][::-1]), runs)


def test_bins_left_equals_searchsorted():
    """The arithmetic binning is np.searchsorted(edges, x) - 1 exactly (super_block.py:114-124),
    also for coordinates ON an edge, outside the network, and for arange's rounded edges."""
    rng = np.random.default_rng(11)
    for mn, siz, n in ((0.5 * 5.12, 5.12, 50), (200.0, 400.0, 49), (0.05, 0.1, 37),
                       (-3.3, 1.0 / 3.0, 20), (1e6 + 0.7, 0.7, 1)):
        e = _edges_sample(mn, siz, n)
        x = np.concatenate([rng.uniform(e[0] - 3 * siz, e[-1] + 3 * siz, 20000), e,
                            np.nextafter(e, -np.inf), np.nextafter(e, np.inf),
                            [e[0] - 1e30, e[-1] + 1e30]])
        assert np.array_equal(bins_left(e, x), np.searchsorted(e, x) - 1)
    e = np.array([0.0, 1.0, 1.5, 7.0, 7.25, 30.0])  # irregular: falls back to the search
    x = rng.uniform(-2, 35, 1000)
    assert np.array_equal(bins_left(e, x), np.searchsorted(e, x) - 1)


def test_no_cpu_fallback_without_gpu():
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    par, data, sec = k3d_configs.config("c2", 0.1)
    k = Krige3d(par, data=data)
    k.verbose = False
    with pytest.raises(_native.K3dError):
        k.kt3d()


def test_slab_ranges():
    assert slab_ranges(256, 8) == [(32 * i, 32 * i + 32) for i in range(8)]
    assert slab_ranges(10, 4) == [(0, 3), (3, 6), (6, 8), (8, 10)]
    assert slab_ranges(1, 2) == [(0, 1), (1, 1)]
    for nz, w in ((79, 8), (5, 8), (256, 3)):
        r = slab_ranges(nz, w)
        assert r[0][0] == 0 and r[-1][1] == nz
        assert all(a[1] == b[0] for a, b in zip(r, r[1:]))


_WORKER = r'''
import os, sys
sys.path.insert(0, %r)
import numpy as np, torch, torch.distributed as dist
import k3d_configs
from oracle import kt3d_oracle as orc
from pygeostatistics_b200.krige3d import slab_ranges, gather_slabs
dist.init_process_group("gloo")
rank, world = dist.get_rank(), dist.get_world_size()
par, data, sec = k3d_configs.config("c2", 0.12)
par["nz"] = par["nz"] - 1                      # odd plane count: unequal slabs
keep = data["z"] < par["nz"]
data = {n: v[keep] for n, v in data.items()}
st = orc.prepare(par, dict(data), ["v"])
plane = par["nx"] * par["ny"]
ranges = slab_ranges(par["nz"], world)
a, b = ranges[rank]
# each rank computes ONLY its slab (the oracle stands in for the kernel on CPU) ...
est, var, _, _ = orc.run(st, node_range=(a * plane, b * plane), neighbours=False)
full = gather_slabs(torch.from_numpy(est), ranges, plane).numpy()
# ... and every rank must end up with the whole grid
ref, _, _, _ = orc.run(st, neighbours=False)
assert full.shape == ref.shape and np.array_equal(full, ref, equal_nan=True), rank
if rank == 0:
    print("GATHER_OK", full.shape[0])
dist.destroy_process_group()
'''


def test_two_rank_z_slab_sharding_gloo(tmp_path):
    script = tmp_path / "worker.py"
    script.write_text(_WORKER % ROOT)
    port = 29500 + os.getpid() % 2000
    out = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1",
                          "--nproc-per-node=2", "--master-addr", "127.0.0.1", "--master-port",
                          str(port), str(script)], capture_output=True, text=True, timeout=300)
    assert out.returncode == 0, out.stderr[-2000:]
    assert "GATHER_OK" in out.stdout


def test_pinned_result_pool_keeps_arrays_private(monkeypatch):
    """Result buffers are pooled page-locked memory; a buffer may only be lent again after
    the ndarray it backed is gone (a second Krige3d must not overwrite the first's results)."""
    import gc
    import torch
    import pygeostatistics_b200.krige3d as k3
    real_empty = torch.empty
    monkeypatch.setattr(torch, "empty", lambda *a, **kw: real_empty(
        *a, **{k: v for k, v in kw.items() if k != "pin_memory"}))
    monkeypatch.setattr(k3, "_PINNED", [])
    _, a1 = k3._pinned_array(10, torch.float64)
    a1[:] = 1.0
    _, a2 = k3._pinned_array(10, torch.float64)
    a2[:] = 2.0
    assert a1[0] == 1.0 and a2[0] == 2.0 and len(k3._PINNED) == 2
    del a1
    gc.collect()
    _, a3 = k3._pinned_array(8, torch.float64)      # reuses the freed buffer
    a3[:] = 3.0
    assert len(k3._PINNED) == 2 and a2[0] == 2.0
    _, a4 = k3._pinned_array(8, torch.uint8)        # other dtype: its own buffer
    assert len(k3._PINNED) == 3 and a4.dtype == np.uint8
    # more than _PINNED_MAX buffers: idle ones of different sizes are dropped (by identity; comparing
    # the entries would compare tensors element-wise), lent ones stay
    keep = [k3._pinned_array(100 + i, torch.float64)[1] for i in range(6)]
    for i in range(20):
        k3._pinned_array(1000 + i, torch.int32)     # result dropped at once: idle
        gc.collect()
    assert len(k3._PINNED) <= max(k3._PINNED_MAX, 9) + 1   # (3 + 6 of them are lent out)
    assert all(any(e[1] is not None and e[1]() is a for e in k3._PINNED) for a in keep)


def test_plane_runs_cut_at_super_block_boundaries():
    """kt3d() pipelines the slab in a few runs of z-planes: they tile [iz0, iz1) exactly and
    every inner cut is the first plane of a super block."""
    from pygeostatistics_b200 import Krige3d

    class _S(object):
        nodez0 = np.array([0, 5, 10, 15, 20, 26, 31, 36, 41, 46, 51, 56, 61, 64])
    k = Krige3d.__new__(Krige3d)
    k.searcher = _S()
    for iz0, iz1 in ((0, 64), (10, 13), (0, 1), (16, 48), (3, 64), (20, 26)):
        runs = k._plane_runs(iz0, iz1)
        assert runs[0][0] == iz0 and runs[-1][1] == iz1 and len(runs) <= 4
        assert all(a[1] == b[0] for a, b in zip(runs[:-1], runs[1:]))
        assert all(a < b for a, b in runs)
        assert all(b in _S.nodez0 for _, b in runs[:-1])


def test_block_of_locations_equals_oracle_getindx():
    """Super block of arbitrary locations (getindx, super_block.py:319-346): host set-up
    and oracle agree, also on edges and outside the network (both clamp)."""
    from oracle import kt3d_oracle as orc
    par, data, sec = k3d_configs.config("c3", 0.2)
    st = orc.prepare(par, dict(data), ['v'])
    k = Krige3d(par, data=data)
    k._preprocess()
    s = SuperBlockSearcher()
    s.nx, s.xmn, s.xsiz, s.ny, s.ymn, s.ysiz = k.nx, k.xmn, k.xsiz, k.ny, k.ymn, k.ysiz
    s.nz, s.zmn, s.zsiz, s.vr, s.MAXSB = k.nz, k.zmn, k.zsiz, k.vr, k.const.MAXSB
    s.setup()
    rng = np.random.default_rng(3)
    m = 5000
    px = rng.uniform(-3, par['nx'] + 3, m); py = rng.uniform(-3, par['ny'] + 3, m)
    pz = rng.uniform(-3, par['nz'] + 3, m)
    edges = st['node_edges'][0]
    px[:len(edges)] = edges
    sb = orc.super_block_of(st, px, py, pz)
    flat = sb[:, 0] + sb[:, 1] * s.nxsup + sb[:, 2] * s.nxsup * s.nysup
    assert np.array_equal(s.block_of_locations(px, py, pz), flat)


def test_krige2d_front_end_host_side(tmp_path):
    """Krige2d (krige2d.py:19-104 interface): parameter file, checks, whitespace-separated
    reader, and the krige3d mapping it runs with -- no GPU needed up to the launch."""
    from pygeostatistics_b200 import Krige2d, _native
    ref = os.path.join(ROOT, "tests", "golden", "krige2d_reference.npz")
    g = np.load(ref)
    par = json.loads(str(g['c1_par'][0]))
    x, y, v = g['c1_xyv']
    data = tmp_path / "d.gslib"
    with open(data, "w") as f:
        f.write("test\n3\nx\ny\npor\n")
        for r in zip(x, y, v):
            f.write("   ".join(repr(float(c)) for c in r) + "\n")
    pf = tmp_path / "k2d.par"
    json.dump(dict(par, datafl=str(data)), open(pf, "w"))
    k = Krige2d(str(pf))
    k.read_data()
    assert k._2d and k.property_name == ['por'] and k.vr.dtype.names == ('x', 'y', 'por', 'z')
    assert np.array_equal(k.vr['x'], x) and np.array_equal(k.vr['por'], v)
    p3 = k.krige3d_params()
    assert (p3['nz'], p3['noct'], p3['ikrige'], p3['radius_hmax']) == (1, 0, par['isk'], float(par['radius']))
    assert p3['ang1'] == [float(a) for a in par['azm']] and p3['aa_hmin'] == [float(a) for a in par['a_min']]
    assert p3['_search_identity'] and p3['_flags'] == _native.FLAG_OK_ONE_SAMPLE
    for bad, exc in ((dict(it=[7]), ValueError), (dict(it=[5]), ValueError),
                     (dict(it=[4], a_max=[1.5]), NotImplementedError),
                     (dict(it=[4], a_max=[2.5]), ValueError), (dict(isk=2), ValueError),
                     (dict(ndmax=65), ValueError)):
        with pytest.raises(exc):
            Krige2d(dict(par, **bad))
    missing = dict(par)
    del missing['radius']
    with pytest.raises(KeyError):
        Krige2d(missing)
    # the krige3d class honours the two front-end switches
    k3 = Krige3d({a: b for a, b in p3.items() if not a.startswith('_')},
                 data={'x': x, 'y': y, 'z': np.zeros_like(x), 'por': v})
    k3.search_identity = True
    k3._preprocess(); k3._set_rotation()
    assert np.array_equal(k3.rotmat[-1], np.eye(3))
    assert k3.mdt == par['isk']


_SHARED_WORKER = r'''
import os, sys
sys.path.insert(0, %r)
import numpy as np, torch, torch.distributed as dist
from pygeostatistics_b200.krige3d import _SharedGrid, slab_ranges
dist.init_process_group("gloo")
rank, world = dist.get_rank(), dist.get_world_size()
plane, nz = 100, 7
spec = (("est", "float64", plane * nz), ("status", "uint8", plane * nz))
sg = _SharedGrid(spec, rank, None)
assert sg.ok and not os.path.exists("/dev/shm/k3d_%%d" %% os.getpid())
base = sg.lend()
a, b = slab_ranges(nz, world)[rank]
sg.view(base, "est")[a * plane: b * plane] = rank + 1.0          # each rank writes ONLY its slab
sg.view(base, "status")[a * plane: b * plane] = rank + 1
dist.barrier()
est, st = sg.view(base, "est"), sg.view(base, "status")           # ... and reads the whole grid
for r, (ra, rb) in enumerate(slab_ranges(nz, world)):
    assert np.all(est[ra * plane: rb * plane] == r + 1.0) and np.all(st[ra * plane: rb * plane] == r + 1)
assert not sg.free()
del base, est, st
assert sg.free()                                                   # lent array gone: reusable
if rank == 0:
    print("SHARED_OK")
dist.destroy_process_group()
'''


def test_shared_result_buffer_two_ranks_gloo(tmp_path):
    """The host buffer the ranks of a node share for the results of a multi-rank kt3d()."""
    script = tmp_path / "worker.py"
    script.write_text(_SHARED_WORKER % ROOT)
    port = 31500 + os.getpid() % 2000
    out = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1",
                          "--nproc-per-node=2", "--master-addr", "127.0.0.1", "--master-port",
                          str(port), str(script)], capture_output=True, text=True, timeout=300)
    assert out.returncode == 0, out.stderr[-2000:]
    assert "SHARED_OK" in out.stdout


def test_data_fingerprint_sees_edits_and_reorderings():
    from pygeostatistics_b200.krige3d import data_fingerprint
    rng = np.random.default_rng(0)
    vr = np.rec.fromarrays([rng.normal(size=50) for _ in range(4)], names='x,y,z,v')
    f0 = data_fingerprint(vr)
    assert data_fingerprint(vr.copy()) == f0
    e = vr.copy(); e['v'][17] = np.nextafter(e['v'][17], 1e9)
    assert data_fingerprint(e) != f0
    p = vr.copy(); p[[3, 9]] = p[[9, 3]]
    assert data_fingerprint(p) != f0
