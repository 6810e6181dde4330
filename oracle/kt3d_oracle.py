"""CPU parity oracle for the krige3d grid-estimation path -- host side.

TEST INFRASTRUCTURE ONLY: imported by tests/, __graft_entry__.smoke() and bench.py's
cpu_baseline / --impl reference legs.  The product (pygeostatistics_b200/) never
imports this module and has its own, separately written, host set-up.

Restates with plain numpy what the reference does once per run before its node loop,
and drives oracle/kt3d_oracle.c for the per-node work:

  read_par            krige3d.py:57-130   (JSON parameter file)
  read_geoeas         gslib_reader.py:25-50
  rotation_matrices   krige3d.py:211-253  (quirk D12 kept as written)
  prepare             krige3d.py:161-209, 255-282, 414-421, 664-721 and
                      super_block.py:87-161 (setup, pickup), :319-346 (getindx)

Pinned against the unmodified reference by tests/test_oracle_golden.py.
"""
import ctypes as C
import json
import os
import subprocess

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
LIB = os.path.join(HERE, "libkt3d_oracle.so")
MAXNST, MAXDRIFT = 4, 9
EPS = np.finfo(float).eps


class OrcParams(C.Structure):
    _fields_ = [
        ("nx", C.c_int32), ("ny", C.c_int32), ("nz", C.c_int32),
        ("xmn", C.c_double), ("ymn", C.c_double), ("zmn", C.c_double),
        ("xsiz", C.c_double), ("ysiz", C.c_double), ("zsiz", C.c_double),
        ("ndmin", C.c_int32), ("ndmax", C.c_int32), ("noct", C.c_int32),
        ("radsqd", C.c_double),
        ("ktype", C.c_int32), ("idrift", C.c_int32 * MAXDRIFT),
        ("itrend", C.c_int32), ("skmean", C.c_double),
        ("tmin", C.c_double), ("tmax", C.c_double),
        ("nst", C.c_int32), ("it", C.c_int32 * MAXNST), ("c0", C.c_double),
        ("cc", C.c_double * MAXNST), ("aa", C.c_double * MAXNST),
        ("rotmat", C.c_double * ((MAXNST + 1) * 9)),
        ("maxcov", C.c_double), ("resc", C.c_double), ("unbias", C.c_double),
        ("cbb", C.c_double), ("bv", C.c_double * MAXDRIFT),
        ("nxsup", C.c_int32), ("nysup", C.c_int32), ("nzsup", C.c_int32),
        ("ndb", C.c_int32), ("flags", C.c_int32),
    ]


_lib = None


def build(force=False):
    """Compile oracle/kt3d_oracle.c -> oracle/libkt3d_oracle.so (gcc, see Makefile)."""
    src = os.path.join(HERE, "kt3d_oracle.c")
    if force or not os.path.exists(LIB) or os.path.getmtime(LIB) < os.path.getmtime(src):
        subprocess.check_call(["make", "-s", "-C", HERE, "-B", "libkt3d_oracle.so"])
    return LIB


def lib():
    global _lib
    if _lib is None:
        build()
        _lib = C.CDLL(LIB)
        assert _lib.kt3d_oracle_sizeof_params() == C.sizeof(OrcParams)
        _lib.orc_cova3.restype = C.c_double
        _lib.orc_cova3.argtypes = [C.POINTER(OrcParams)] + [C.c_double] * 6
        _lib.orc_block_covariance.restype = C.c_double
    return _lib


def _p(a, t):
    return a.ctypes.data_as(C.POINTER(t)) if a is not None else None


def read_par(path):
    with open(path, "r") as f:
        return json.load(f)


def read_geoeas(path):
    """gslib_reader.py:25-50: title, ncols, names, tab separated rows; z=0 appended if absent."""
    with open(path, "r") as f:
        f.readline()
        ncols = int(f.readline().strip())
        names = [f.readline().strip() for _ in range(ncols)]
        rows = [ln.split("\t") for ln in f if ln.strip()]
    data = np.array(rows, dtype=np.float64).reshape(-1, ncols)
    cols = {n: np.ascontiguousarray(data[:, i]) for i, n in enumerate(names)}
    if "z" not in cols:
        cols["z"] = np.zeros(data.shape[0])
        names = names + ["z"]
    return names, cols


def rotation_matrices(ang1, ang2, ang3, anis1, anis2):
    """krige3d.py:211-253.  `ang <= 0 and ang < 270` is the reference's test (D12)."""
    ang1 = np.asarray(ang1); ang2 = np.asarray(ang2); ang3 = np.asarray(ang3)
    alpha = np.array([np.deg2rad(90 - a) if (a <= 0 and a < 270) else
                      np.deg2rad(450 - a) for a in ang1])
    beta = np.deg2rad(-ang2)
    theta = np.deg2rad(ang3)
    sina, sinb, sint = np.sin(alpha), np.sin(beta), np.sin(theta)
    cosa, cosb, cost = np.cos(alpha), np.cos(beta), np.cos(theta)
    afac1 = 1.0 / np.maximum(anis1, EPS)
    afac2 = 1.0 / np.maximum(anis2, EPS)
    r = np.empty((len(ang1), 3, 3))
    r[:, 0, 0] = cosb * cosa
    r[:, 0, 1] = cosb * sina
    r[:, 0, 2] = -sinb
    r[:, 1, 0] = afac1 * (-cost * sina + sint * sinb * cosa)
    r[:, 1, 1] = afac1 * (cost * cosa + sint * sinb * sina)
    r[:, 1, 2] = afac1 * (sint * cosb)
    r[:, 2, 0] = afac2 * (sint * sina + cost * sinb * cosa)
    r[:, 2, 1] = afac2 * (-sint * cosa + cost * sinb * sina)
    r[:, 2, 2] = afac2 * (cost * cosb)
    return r


def _maxsb(n):
    """krige3d.py:167-181"""
    if n <= 1:
        return 1
    return min(int(n / 2), 50)


def _edges_numpy(mn, siz, n):
    """super_block.py:111-113 -- numpy's arange."""
    return np.arange(mn - 0.5 * siz, mn + (n + 1) * siz + 1, siz)


def _edges_numba(mn, siz, n):
    """super_block.py:331-333 -- Numba's arange: start + i*step, ceil((stop-start)/step) items."""
    start = mn - 0.5 * siz
    stop = mn + (n + 1) * siz + 1
    cnt = int(np.ceil((stop - start) / siz))
    return start + np.arange(cnt) * siz


def prepare(par, cols, prop_names):
    """Everything the reference computes once before its node loop."""
    st = {}
    nx, ny, nz = par['nx'], par['ny'], par['nz']
    nst = par['nst']
    radsqd = par['radius_hmax'] * par['radius_hmax']
    sanis1 = par['radius_hmin'] / par['radius_hmax']
    sanis2 = par['radius_vert'] / par['radius_hmax']
    anis1 = np.array(par['aa_hmin']) / np.maximum(par['aa_hmax'], EPS)
    anis2 = np.array(par['aa_vert']) / np.maximum(par['aa_hmax'], EPS)
    rot = rotation_matrices(np.append(par['ang1'], par['sang1']),
                            np.append(par['ang2'], par['sang2']),
                            np.append(par['ang3'], par['sang3']),
                            np.append(anis1, sanis1), np.append(anis2, sanis2))
    if par.get('_search_identity'):   # krige2d front end: plain dx^2 + dy^2 search distance
        rot[-1] = np.eye(3)
    maxcov = par['c0']
    for i in range(nst):
        maxcov += 999 if par['it'][i] == 4 else par['cc'][i]
    if radsqd < 1:
        resc = 2 * par['radius_hmax'] / max(maxcov, 0.0001)
    else:
        resc = (4 * radsqd) / max(maxcov, 0.0001)
    if resc <= 0:
        raise ValueError("rescaling value is wrong, {}".format(resc))
    resc = 1 / resc
    # super blocks (super_block.py:98-108)
    nxsup, nysup, nzsup = min(nx, _maxsb(nx)), min(ny, _maxsb(ny)), min(nz, _maxsb(nz))
    xsizsup = nx * par['xsiz'] / nxsup
    ysizsup = ny * par['ysiz'] / nysup
    zsizsup = nz * par['zsiz'] / nzsup
    xmnsup = (par['xmn'] - 0.5 * par['xsiz']) + 0.5 * xsizsup
    ymnsup = (par['ymn'] - 0.5 * par['ysiz']) + 0.5 * ysizsup
    zmnsup = (par['zmn'] - 0.5 * par['zsiz']) + 0.5 * zsizsup
    x, y, z = cols['x'], cols['y'], cols['z']
    bx = np.searchsorted(_edges_numpy(xmnsup, xsizsup, nxsup), x) - 1
    by = np.searchsorted(_edges_numpy(ymnsup, ysizsup, nysup), y) - 1
    bz = np.searchsorted(_edges_numpy(zmnsup, zsizsup, nzsup), z) - 1
    blk = bx + by * nxsup + bz * nxsup * nysup
    nsup = nxsup * nysup * nzsup
    if blk.size and (blk.min() < 0 or blk.max() >= nsup):
        raise ValueError("sample outside the super block network (D13)")
    sort_index = np.argsort(blk, kind='stable')            # D14: stable
    nisb = np.cumsum(np.bincount(blk, minlength=nsup)).astype(np.int64)
    # node -> super block LUTs, Numba-flavour edges (super_block.py:319-346)
    ex, ey, ez = (_edges_numba(xmnsup, xsizsup, nxsup),
                  _edges_numba(ymnsup, ysizsup, nysup),
                  _edges_numba(zmnsup, zsizsup, nzsup))
    lutx = (np.searchsorted(ex, par['xmn'] + np.arange(nx) * par['xsiz']) - 1).astype(np.int32)
    luty = (np.searchsorted(ey, par['ymn'] + np.arange(ny) * par['ysiz']) - 1).astype(np.int32)
    lutz = (np.searchsorted(ez, par['zmn'] + np.arange(nz) * par['zsiz']) - 1).astype(np.int32)
    # template (super_block.py:226-265)
    L = lib()
    rs = np.ascontiguousarray(rot[-1].ravel())
    L.orc_pickup.restype = C.c_int
    args = [C.c_int(nxsup), C.c_int(nysup), C.c_int(nzsup), C.c_double(xsizsup),
            C.c_double(ysizsup), C.c_double(zsizsup), _p(rs, C.c_double),
            C.c_double(radsqd)]
    nt = L.orc_pickup(*args, None, None, None)
    tx = np.zeros(max(nt, 1), np.int32); ty = tx.copy(); tz = tx.copy()
    L.orc_pickup(*args, _p(tx, C.c_int32), _p(ty, C.c_int32), _p(tz, C.c_int32))
    # block discretisation (krige3d.py:692-707, repair D9: tensor product, x outer)
    nxdis, nydis, nzdis = (max(1, par['nxdis']), max(1, par['nydis']), max(1, par['nzdis']))
    xdis, ydis, zdis = par['xsiz'] / nxdis, par['ysiz'] / nydis, par['zsiz'] / nzdis
    ax = np.arange(0, nxdis, 1) * xdis + (-0.5 * par['xsiz'] + 0.5 * xdis)
    ay = np.arange(0, nydis, 1) * ydis + (-0.5 * par['ysiz'] + 0.5 * ydis)
    az = np.arange(0, nzdis, 1) * zdis + (-0.5 * par['zsiz'] + 0.5 * zdis)
    gx, gy, gz = np.meshgrid(ax, ay, az, indexing='ij')
    xdb, ydb, zdb = (np.ascontiguousarray(g.ravel()) for g in (gx, gy, gz))
    ndb = nxdis * nydis * nzdis
    bv = np.array([np.mean(xdb), np.mean(ydb), np.mean(zdb), np.mean(xdb * xdb),
                   np.mean(ydb * ydb), np.mean(zdb * zdb), np.mean(xdb * ydb),
                   np.mean(xdb * zdb), np.mean(ydb * zdb)]) * resc
    ktype = par['ikrige']
    idrift = [bool(b) for b in par['idrift']]
    if ktype in (0, 2):
        idrift = [False] * 9
    P = OrcParams()
    P.nx, P.ny, P.nz = nx, ny, nz
    P.xmn, P.ymn, P.zmn = par['xmn'], par['ymn'], par['zmn']
    P.xsiz, P.ysiz, P.zsiz = par['xsiz'], par['ysiz'], par['zsiz']
    P.ndmin, P.ndmax, P.noct = par['ndmin'], par['ndmax'], par['noct']
    P.radsqd = radsqd
    P.ktype = ktype
    for i in range(9):
        P.idrift[i] = int(idrift[i])
    P.itrend = int(bool(par['itrend']))
    P.skmean = par['skmean']
    P.tmin, P.tmax = par['tmin'], par['tmax']
    P.nst = nst
    for i in range(nst):
        P.it[i] = int(par['it'][i]); P.cc[i] = par['cc'][i]; P.aa[i] = par['aa_hmax'][i]
    P.c0 = par['c0']
    for i, v in enumerate(rot.ravel()):
        P.rotmat[i] = v
    P.maxcov, P.resc, P.unbias = maxcov, resc, maxcov
    for i in range(9):
        P.bv[i] = bv[i]
    P.nxsup, P.nysup, P.nzsup = nxsup, nysup, nzsup
    P.ndb = ndb
    P.flags = int(par.get('_flags', 0))
    L.orc_block_covariance.argtypes = [C.POINTER(OrcParams)] + [C.POINTER(C.c_double)] * 3
    P.cbb = maxcov
    P.cbb = L.orc_block_covariance(C.byref(P), _p(xdb, C.c_double), _p(ydb, C.c_double),
                                   _p(zdb, C.c_double))
    sec = cols[prop_names[1]] if len(prop_names) > 1 else np.zeros_like(x)
    st.update(P=P, rotmat=rot, maxcov=maxcov, resc=resc, radsqd=radsqd, bv=bv,
              sort_index=sort_index, nisb=nisb, lutx=lutx, luty=luty, lutz=lutz,
              tx=tx[:nt], ty=ty[:nt], tz=tz[:nt], ntmpl=nt, xdb=xdb, ydb=ydb, zdb=zdb,
              sx=np.ascontiguousarray(x[sort_index]), sy=np.ascontiguousarray(y[sort_index]),
              sz=np.ascontiguousarray(z[sort_index]),
              sv=np.ascontiguousarray(cols[prop_names[0]][sort_index]),
              ssec=np.ascontiguousarray(sec[sort_index]),
              nsup=(nxsup, nysup, nzsup), sizsup=(xsizsup, ysizsup, zsizsup),
              mnsup=(xmnsup, ymnsup, zmnsup), block_of_sample=blk, cbb=P.cbb,
              node_edges=(ex, ey, ez))
    return st


def run(st, extgrid=None, node_range=None, node_list=None, threads=1, neighbours=True):
    """Krige a node range / list with the C oracle.  Returns est, var, nbr_idx, nbr_cnt
    (neighbour indices are positions in the super-block-sorted sample arrays)."""
    P = st['P']
    L = lib()
    if node_list is not None:
        node_list = np.ascontiguousarray(node_list, dtype=np.int64)
        count, n0, n1 = len(node_list), 0, 0
    else:
        n0, n1 = node_range if node_range is not None else (0, P.nx * P.ny * P.nz)
        count = n1 - n0
    est = np.full(count, np.nan); var = np.full(count, np.nan)
    nbr_idx = np.full((count, P.ndmax), -1, np.int32) if neighbours else None
    nbr_cnt = np.zeros(count, np.int32)
    if P.ktype in (2, 3):
        if extgrid is None:
            raise ValueError("ikrige 2/3 needs the external drift grid")
        extgrid = np.ascontiguousarray(extgrid, dtype=np.float64)
    D, I32, I64 = C.c_double, C.c_int32, C.c_int64
    L.kt3d_oracle_run.restype = C.c_int
    L.kt3d_oracle_run(
        C.byref(P), C.c_int64(len(st['sx'])), _p(st['sx'], D), _p(st['sy'], D),
        _p(st['sz'], D), _p(st['sv'], D), _p(st['ssec'], D), _p(st['nisb'], I64),
        C.c_int(st['ntmpl']), _p(st['tx'], I32), _p(st['ty'], I32), _p(st['tz'], I32),
        _p(st['lutx'], I32), _p(st['luty'], I32), _p(st['lutz'], I32),
        _p(extgrid, D), _p(st['xdb'], D), _p(st['ydb'], D), _p(st['zdb'], D),
        C.c_int64(n0), C.c_int64(n1), _p(node_list, I64), C.c_int64(count),
        _p(est, D), _p(var, D), _p(nbr_idx, I32), _p(nbr_cnt, I32), C.c_int(threads))
    return est, var, nbr_idx, nbr_cnt


def super_block_of(st, px, py, pz):
    """getindx (super_block.py:319-346): super block of arbitrary locations, with the
    Numba-flavoured edges; clamped to the network (the reference does not clamp: parity is
    defined for locations inside the grid extent)."""
    out = []
    for e, c, n in zip(st['node_edges'], (px, py, pz), st['nsup']):
        out.append(np.clip(np.searchsorted(e, np.asarray(c, dtype=np.float64)) - 1, 0, n - 1))
    return np.ascontiguousarray(np.stack(out, axis=1).astype(np.int32))


def run_points(st, px, py, pz, ext=None, koption=1, threads=1, neighbours=True):
    """Cross-validation (koption 1) / jackknife (koption 2) at the listed locations
    (krige3d.py:310-332, 359-363).  Returns est, var, nbr_idx, nbr_cnt indexed by location."""
    P = st['P']
    L = lib()
    D, I32, I64 = C.c_double, C.c_int32, C.c_int64
    px, py, pz = (np.ascontiguousarray(a, dtype=np.float64) for a in (px, py, pz))
    n = len(px)
    sb = super_block_of(st, px, py, pz)
    if P.ktype in (2, 3):
        if ext is None:
            raise ValueError("ikrige 2/3 needs the external value of every location")
        ext = np.ascontiguousarray(ext, dtype=np.float64)
    else:
        ext = None
    est = np.full(n, np.nan); var = np.full(n, np.nan)
    nbr_idx = np.full((n, P.ndmax), -1, np.int32) if neighbours else None
    nbr_cnt = np.zeros(n, np.int32)
    L.kt3d_oracle_run_points.restype = C.c_int
    L.kt3d_oracle_run_points(
        C.byref(P), C.c_int64(len(st['sx'])), _p(st['sx'], D), _p(st['sy'], D),
        _p(st['sz'], D), _p(st['sv'], D), _p(st['ssec'], D), _p(st['nisb'], I64),
        C.c_int(st['ntmpl']), _p(st['tx'], I32), _p(st['ty'], I32), _p(st['tz'], I32),
        _p(st['xdb'], D), _p(st['ydb'], D), _p(st['zdb'], D), C.c_int64(n),
        _p(px, D), _p(py, D), _p(pz, D), _p(ext, D), _p(sb, I32), C.c_int(int(koption)),
        _p(est, D), _p(var, D), _p(nbr_idx, I32), _p(nbr_cnt, I32), C.c_int(threads))
    return est, var, nbr_idx, nbr_cnt


def krige_files(par_path, threads=1, **kw):
    """Krige3d(par).kt3d() equivalent driven from files (datafl relative to CWD)."""
    par = read_par(par_path)
    names, cols = read_geoeas(par['datafl'])
    prop = [n for n in names if n not in ('x', 'y', 'z')]
    st = prepare(par, cols, prop)
    ext = None
    if par['ikrige'] in (2, 3):
        en, ec = read_geoeas(par['secfl'])
        ext = ec[en[par['iseccol']]] if isinstance(par['iseccol'], int) else ec[par['iseccol']]
    return st, run(st, extgrid=ext, threads=threads, **kw)
