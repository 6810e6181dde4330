"""CPU model of the search kernel's algorithm (k3d_search_kernel) against
the oracle's neighbour lists.

The kernel keeps exactly what the reference keeps only if three things hold that have nothing
to do with the hardware: (1) dropping the candidates no node of the tile can reach, (2)
streaming the rest in order of their (binned) distance from the tile centre and STOPPING by the
triangle inequality, (3) the per-octant / ndmax bookkeeping with the reference's tie rule.  This
restates those steps in numpy -- same bins, same chunks of 16, same bounds and margins -- and
checks the ordered lists it produces against oracle/ on small cases (test infrastructure; the
product has no CPU path)."""
import numpy as np
import pytest

import k3d_configs
from oracle import kt3d_oracle as orc

SBINS, SCHUNK = 64, 16


def sqdist_exact(dx, dy, dz, r):
    """super_block.py:349-378, the association the kernel uses (no fused multiply-add)."""
    c0 = (r[0] * dx + r[1] * dy) + r[2] * dz
    c1 = (r[3] * dx + r[4] * dy) + r[5] * dz
    c2 = (r[6] * dx + r[7] * dy) + r[8] * dz
    return (c0 * c0 + c1 * c1) + c2 * c2


def octant_of(dx, dy, dz):
    """sgsim.py:945-964 (repair D5) on sample - node."""
    lower = dz < 0
    iq = np.full(dx.shape, 3)
    iq[(dx <= 0) & (dy > 0)] = 0
    iq[(dx > 0) & (dy >= 0)] = 1
    iq[(dx < 0) & (dy <= 0)] = 2
    return iq + 4 * lower


def model_search(st, par):
    """Neighbour lists of every grid node by the kernel's algorithm; also the number of
    candidates each node tested."""
    nx, ny, nz = par['nx'], par['ny'], par['nz']
    nxs, nys, nzs = st['nsup']
    rs = np.asarray(st['rotmat'][-1], dtype=np.float64).ravel()
    radsqd, ndmax, noct = st['radsqd'], par['ndmax'], par['noct']
    sx, sy, sz = st['sx'], st['sy'], st['sz']
    start = np.concatenate([[0], st['nisb']]).astype(np.int64)
    lutx, luty, lutz = (np.asarray(a) for a in (st['lutx'], st['luty'], st['lutz']))
    nbr = np.full((nx * ny * nz, ndmax), -1, np.int64)
    cnt = np.zeros(nx * ny * nz, np.int64)
    tested = np.zeros(nx * ny * nz, np.int64)
    for bz in range(nzs):
        izs = np.nonzero(lutz == bz)[0]
        for by in range(nys):
            iys = np.nonzero(luty == by)[0]
            for bx in range(nxs):
                ixs = np.nonzero(lutx == bx)[0]
                if not (len(ixs) and len(iys) and len(izs)):
                    continue
                # 1. candidates of the tile, ascending sample index
                tb = np.stack([bx + st['tx'], by + st['ty'], bz + st['tz']], axis=1)
                ok = (tb[:, 0] >= 0) & (tb[:, 0] < nxs) & (tb[:, 1] >= 0) & (tb[:, 1] < nys) & \
                     (tb[:, 2] >= 0) & (tb[:, 2] < nzs)
                flat = np.sort(tb[ok, 0] + nxs * (tb[ok, 1] + nys * tb[ok, 2]))
                cand = np.concatenate([np.arange(start[b], start[b + 1]) for b in flat]) \
                    if len(flat) else np.zeros(0, np.int64)
                # 2. centre, reach, bins, stable order
                cx = par['xmn'] + (ixs[0] + 0.5 * (len(ixs) - 1)) * par['xsiz']
                cy = par['ymn'] + (iys[0] + 0.5 * (len(iys) - 1)) * par['ysiz']
                cz = par['zmn'] + (izs[0] + 0.5 * (len(izs) - 1)) * par['zsiz']
                ex, ey, ez = (0.5 * (len(a) - 1) * abs(s) for a, s in
                              ((ixs, par['xsiz']), (iys, par['ysiz']), (izs, par['zsiz'])))
                rho2 = max(sqdist_exact(ex, sy_ * ey, sz_ * ez, rs) for sy_ in (1, -1) for sz_ in (1, -1))
                reach = (np.sqrt(radsqd) + np.sqrt(rho2)) * (1.0 + 1e-9)
                hcmax = reach * reach
                hc = sqdist_exact(cx - sx[cand], cy - sy[cand], cz - sz[cand], rs)
                keep = ~(hc > hcmax)
                cand, hc = cand[keep], hc[keep]
                bins = np.minimum(SBINS - 1, (hc * (SBINS / hcmax)).astype(np.int64))
                order = np.argsort(bins, kind='stable')
                cand, bins = cand[order], bins[order]
                binw = hcmax / SBINS * (1.0 - 1e-12)
                # 3. one "thread" per node
                for iz in izs:
                    for iy in iys:
                        for ix in ixs:
                            node = ix + nx * (iy + ny * iz)
                            x = par['xmn'] + ix * par['xsiz']
                            y = par['ymn'] + iy * par['ysiz']
                            z = par['zmn'] + iz * par['zsiz']
                            rho = np.sqrt(sqdist_exact(x - cx, y - cy, z - cz, rs))
                            dx, dy, dz = x - sx[cand], y - sy[cand], z - sz[cand]
                            h = sqdist_exact(dx, dy, dz, rs)
                            oc = octant_of(-dx, -dy, -dz) if noct > 0 else np.zeros(len(cand), np.int64)
                            nl, cap = (8, noct) if noct > 0 else (1, ndmax)
                            lists = [[] for _ in range(nl)]   # sorted (h, index) per list
                            c0 = 0
                            while c0 < len(cand):
                                bound = max((l[-1][0] if len(l) == cap else np.inf) for l in lists)
                                s = np.sqrt(min(bound, radsqd)) + rho
                                if bins[c0] * binw > s * s * (1.0 + 1e-9):
                                    break                      # nothing behind can be kept
                                c1 = min(c0 + SCHUNK, len(cand))
                                for c in range(c0, c1):
                                    if h[c] > radsqd:
                                        continue
                                    l = lists[oc[c]]
                                    key = (h[c], cand[c])
                                    if len(l) == cap:
                                        if key > l[-1]:
                                            continue
                                        l.pop()
                                    l.append(key)
                                    l.sort()
                                tested[node] += c1 - c0
                                c0 = c1
                            kept = sorted(k for l in lists for k in l)[:ndmax]
                            cnt[node] = len(kept)
                            nbr[node, :len(kept)] = [k[1] for k in kept]
    return nbr, cnt, tested


@pytest.mark.parametrize("name,scale,upd", [
    ("c3", 0.1, {}),                                     # octant search, anisotropic rotated ellipsoid
    ("c3", 0.1, dict(noct=1, ndmax=6)),                  # ndmax cuts what the octants keep
    ("c2", 0.15, {}),                                    # no octants, ndmax 16
    ("c4", 0.09, dict(ndmax=12)),                        # no octants, isotropic
])
def test_search_algorithm_keeps_what_the_reference_keeps(name, scale, upd):
    par, data, sec = k3d_configs.config(name, scale)
    par = dict(par, **upd)
    st = orc.prepare(par, dict(data), ['v'])
    _, _, onbr, ocnt = orc.run(st, threads=4)
    nbr, cnt, tested = model_search(st, par)
    assert np.array_equal(cnt, ocnt)
    assert np.array_equal(nbr, onbr.astype(np.int64))
    # the cases are not trivial, and the early stop does something
    assert ocnt.max() >= min(par['ndmax'], 8) and tested.mean() > 0


def test_ties_on_distance_go_to_the_lower_index():
    """Samples mirrored about a node are at exactly the same distance: the lower sorted index
    must win, in the octant lists and at the ndmax cut."""
    par = dict(k3d_configs.BASE)
    par.update(nx=9, ny=9, nz=3, ikrige=0, ndmax=3, radius_hmax=4.0, radius_hmin=4.0, radius_vert=4.0)
    base = np.array([[4.5, 4.5, 1.5]])
    offs = np.array([[1, 0, 0], [-1, 0, 0], [0, 1, 0], [0, -1, 0], [0, 0, 1], [0, 0, -1],
                     [2, 0, 0], [-2, 0, 0]], dtype=float)
    p = base + offs
    data = {'x': p[:, 0].copy(), 'y': p[:, 1].copy(), 'z': p[:, 2].copy(), 'v': np.arange(8.0)}
    for noct in (0, 1):
        pp = dict(par, noct=noct)
        st = orc.prepare(pp, dict(data), ['v'])
        _, _, onbr, ocnt = orc.run(st, threads=1)
        nbr, cnt, _ = model_search(st, pp)
        assert np.array_equal(cnt, ocnt) and np.array_equal(nbr, onbr.astype(np.int64))
