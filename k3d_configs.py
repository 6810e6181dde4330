"""The benchmark / parity configurations of BASELINE.json, as
in-memory parameter mappings + synthetic sample columns.  Shared by tests/ and bench.py.

All synthetic grids use xsiz=ysiz=zsiz=1, xmn=ymn=zmn=0.5 (extent (0,n) per axis); samples
are numpy.random.default_rng(seed).uniform strictly inside the extent, de-duplicated.
`scale` shrinks the grid edge and the sample count together (density is preserved) so the
CPU oracle can check the same configuration in seconds.
"""
import json
import os

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))

BASE = {
    'datafl': 'data.gslib', 'icolx': 0, 'icoly': 1, 'icolz': 2, 'icolvr': 3, 'icolsec': 4,
    'tmin': -1.0e21, 'tmax': 1.0e21, 'option': 0, 'jackfl': 'jackfl.dat', 'jicolx': 0,
    'jicoly': 1, 'jicolz': 2, 'jicolvr': 3, 'jicolsec': 4, 'idbg': 0, 'dbgfl': 'kt3d.dbg',
    'outfl': 'out.dat', 'nx': 16, 'xmn': 0.5, 'xsiz': 1.0, 'ny': 16, 'ymn': 0.5, 'ysiz': 1.0,
    'nz': 16, 'zmn': 0.5, 'zsiz': 1.0, 'nxdis': 1, 'nydis': 1, 'nzdis': 1, 'ndmin': 1,
    'ndmax': 16, 'noct': 0, 'radius_hmax': 14.0, 'radius_hmin': 14.0, 'radius_vert': 14.0,
    'sang1': 0.0, 'sang2': 0.0, 'sang3': 0.0, 'ikrige': 0, 'skmean': 0.0,
    'idrift': [False] * 9, 'itrend': False, 'secfl': 'secfl.dat', 'iseccol': 0, 'nst': 1,
    'c0': 0.1, 'it': [1], 'cc': [0.9], 'ang1': [0.0], 'ang2': [0.0], 'ang3': [0.0],
    'aa_hmax': [24.0], 'aa_hmin': [24.0], 'aa_vert': [24.0],
}

NESTED = dict(nst=2, it=[2, 1], c0=0.1, cc=[0.4, 0.5], ang1=[30.0, 60.0], ang2=[10.0, 0.0],
              ang3=[0.0, 0.0], aa_hmax=[40.0, 80.0], aa_hmin=[20.0, 50.0],
              aa_vert=[10.0, 25.0])


def external_field(x, y, z):
    """c5's smooth positive secondary variable e(x,y,z)."""
    return 10.0 + 0.01 * x + 0.02 * y + 0.05 * z + np.sin(x / 40.0)


def _samples(seed, n, ext):
    rng = np.random.default_rng(seed)
    pts = rng.uniform(0.0, 1.0, (n, 3)) * np.array(ext, dtype=float)
    pts = np.unique(pts, axis=0)
    rng.shuffle(pts)
    return rng, pts[:, 0].copy(), pts[:, 1].copy(), pts[:, 2].copy()


def _scaled(n, scale, lo=4):
    return max(lo, int(round(n * scale)))


def config(name, scale=1.0):
    """Returns (par, data, secondary): the parameter mapping, the ordered sample columns
    and the per-node external-drift array (or None)."""
    par = dict(BASE)
    sec = None
    if name == 'c1':
        par = json.load(open(os.path.join(HERE, 'tests', 'golden', 'c1_test_krige3d.par')))
        par['datafl'] = os.path.join(HERE, 'tests', 'golden', 'c1_test.gslib')
        return par, None, None
    if name == 'c1ok':
        par, _, _ = config('c1')
        par['ikrige'] = 1
        return par, None, None
    if name == 'c2':      # 10k samples, 128^3, SK, spherical + nugget, ndmax 16, isotropic
        n = _scaled(128, scale)
        nd = max(30, int(round(10000 * scale ** 3)))
        par.update(nx=n, ny=n, nz=n, ikrige=0, skmean=0.0, ndmax=16)
        rng, x, y, z = _samples(2, nd, (n, n, n))
        v = rng.normal(size=len(x))
    elif name in ('c3', 'c4'):
        n = _scaled(256, scale)
        nd = max(60, int(round(100000 * scale ** 3)))
        par.update(NESTED)
        par.update(nx=n, ny=n, nz=n, ikrige=1, ndmax=32)
        if name == 'c3':  # OK, nested anisotropic exp + sph, octant search
            par.update(radius_hmax=24.0, radius_hmin=16.0, radius_vert=12.0, sang1=30.0, noct=4)
            rng, x, y, z = _samples(3, nd, (n, n, n))
            v = rng.normal(size=len(x))
        else:             # UK with all nine drift terms
            par.update(radius_hmax=16.0, radius_hmin=16.0, radius_vert=16.0, noct=0,
                       idrift=[True] * 9)
            rng, x, y, z = _samples(4, nd, (n, n, n))
            v = (1.0 + 0.02 * x - 0.01 * y + 0.03 * z + 1e-4 * (x * x - y * y + 0.5 * z * z)
                 + 2e-4 * (x * y - x * z + y * z) + rng.normal(size=len(x)))
    elif name == 'c5':    # 1M samples, 512x512x256, block kriging 4x4x2 + external drift
        nx = _scaled(512, scale); nz = _scaled(256, scale)
        nd = max(60, int(round(1000000 * scale ** 3)))
        par.update(NESTED)
        par.update(nx=nx, ny=nx, nz=nz, ikrige=3, ndmax=32, nxdis=4, nydis=4, nzdis=2,
                   radius_hmax=12.0, radius_hmin=12.0, radius_vert=12.0, noct=0, iseccol=0)
        rng, x, y, z = _samples(5, nd, (nx, nx, nz))
        e = external_field(x, y, z)
        v = 0.3 * e + rng.normal(size=len(x))
        gz, gy, gx = np.meshgrid(0.5 + np.arange(nz), 0.5 + np.arange(nx), 0.5 + np.arange(nx),
                                 indexing='ij')
        sec = external_field(gx, gy, gz).ravel()
        return par, {'x': x, 'y': y, 'z': z, 'v': v, 'sec': e}, sec
    else:
        raise KeyError(name)
    return par, {'x': x, 'y': y, 'z': z, 'v': v}, sec


def write_files(dirname, par, data, sec):
    """Writes data.gslib / secfl.dat / case.par into dirname (tab separated, repr precision);
    returns the .par path.  The par's datafl/secfl become absolute paths."""
    par = dict(par)
    if data is not None:
        names = list(data.keys())
        path = os.path.join(dirname, 'data.gslib')
        with open(path, 'w') as f:
            f.write('synthetic\n%d\n' % len(names))
            f.write('\n'.join(names) + '\n')
            for row in zip(*[data[n] for n in names]):
                f.write('\t'.join(repr(float(v)) for v in row) + '\n')
        par['datafl'] = path
    if sec is not None:
        path = os.path.join(dirname, 'secfl.dat')
        with open(path, 'w') as f:
            f.write('secondary\n1\nsec\n')
            f.write('\n'.join(repr(float(v)) for v in sec) + '\n')
        par['secfl'] = path
    pfile = os.path.join(dirname, 'case.par')
    with open(pfile, 'w') as f:
        f.write(json.dumps(par, sort_keys=True, indent=4))
    return pfile
