"""Super-block index for the GPU neighbour search.

Replaces SuperBlockSearcher.setup()/pickup() (reference super_block.py:87-161) with a
form the CUDA kernel consumes directly:

  sbstart   exclusive prefix of the per-super-block sample counts (int32, nsup+1)
  runs      the search template as x-contiguous runs of super-block offsets, so that
            every run maps to ONE contiguous range of the sorted sample arrays
  node?0    per axis, first grid node of each super-block column (inverse of the
            node -> super-block lookup, super_block.py:319-346)

The once-per-run work that decides parity is done with the same numpy calls the
reference makes (np.arange / np.searchsorted edges, super_block.py:111-124); the
template itself is evaluated on the GPU by k3d_build_template.
"""
import ctypes as C

import numpy as np
import torch

from . import _native


def _edges_sample(mn, siz, n):
    # super_block.py:111-113: numpy's arange
    return np.arange(mn - 0.5 * siz, mn + (n + 1) * siz + 1, siz)


def bins_left(edges, x):
    """np.searchsorted(edges, x, side='left') - 1 for equally spaced `edges`, i.e. the last
    edge strictly below x (-1 when there is none): an arithmetic guess corrected against
    the edge values themselves, so the result is the reference's (super_block.py:114-124)
    bit for bit at a third of the cost of the binary search (finite x only)."""
    x = np.ascontiguousarray(x, dtype=np.float64)
    last = len(edges) - 1
    step = (edges[-1] - edges[0]) / last
    with np.errstate(invalid='ignore', over='ignore'):
        g = np.floor((x - edges[0]) / step)
    g = np.clip(np.nan_to_num(g, nan=0.0), 0, last).astype(np.int64)
    for _ in range(3):
        down = (edges[g] >= x) & (g > 0)
        up = ~down & (g < last) & (edges[np.minimum(g + 1, last)] < x)
        if not (down.any() or up.any()):
            return np.where(edges[g] >= x, -1, g)
        g = g - down + up
    return np.searchsorted(edges, x) - 1  # irregular edges: the binary search itself


def _edges_node(mn, siz, n):
    # super_block.py:331-333 runs under Numba, whose arange is start + i*step with
    # ceil((stop-start)/step) items; it can differ from numpy's in the last ulp.
    start = mn - 0.5 * siz
    stop = mn + (n + 1) * siz + 1
    count = int(np.ceil((stop - start) / siz))
    return start + np.arange(count) * siz


class SuperBlockSearcher(object):
    """Attribute-compatible with the reference's SuperBlockSearcher after setup() and
    pickup(): nxsup.., xsizsup.., xmnsup.., nisb (inclusive cumulative counts),
    sort_index, nsbtosr, ixsbtosr/iysbtosr/izsbtosr."""

    def __init__(self):
        self.nx = self.xmn = self.xsiz = None
        self.ny = self.ymn = self.ysiz = None
        self.nz = self.zmn = self.zsiz = None
        self.vr = None
        self.columns = None
        self.MAXSB = []
        self.rotmat = None
        self.radsqd = None
        self.noct = None
        self.nisb = None
        self.sort_index = None
        self.nsbtosr = None
        self.ixsbtosr = self.iysbtosr = self.izsbtosr = None
        # GPU-facing products
        self.sbstart = None
        self.runs = None
        self.nodex0 = self.nodey0 = self.nodez0 = None
        self.block_of_sample = None

    def setup(self, device=None):
        """Geometry + binning + stable sort (super_block.py:87-138).  With a CUDA `device`
        the stable sort of the block ids runs there (same permutation: a stable sort is
        unique), which is ten times quicker than numpy's for 10^5..10^6 samples."""
        self.nxsup = min(self.nx, self.MAXSB[0])
        self.nysup = min(self.ny, self.MAXSB[1])
        self.nzsup = min(self.nz, self.MAXSB[2])
        self.xsizsup = self.nx * self.xsiz / self.nxsup
        self.ysizsup = self.ny * self.ysiz / self.nysup
        self.zsizsup = self.nz * self.zsiz / self.nzsup
        self.xmnsup = (self.xmn - 0.5 * self.xsiz) + 0.5 * self.xsizsup
        self.ymnsup = (self.ymn - 0.5 * self.ysiz) + 0.5 * self.ysizsup
        self.zmnsup = (self.zmn - 0.5 * self.zsiz) + 0.5 * self.zsizsup
        nsup = self.nxsup * self.nysup * self.nzsup

        def bins(mn, siz, n, coord):
            b = bins_left(_edges_sample(mn, siz, n), coord)
            # samples outside the network go to the nearest edge block (the behaviour
            # the reference documents at super_block.py:19-20 but does not implement)
            return np.clip(b, 0, n - 1)

        if getattr(self, 'columns', None) is not None:  # name -> contiguous column, file order
            names = tuple(self.columns.keys())
            cols = dict(self.columns)
            rec_dtype = np.dtype({'names': list(names), 'formats': ['f8'] * len(names)})
        else:
            names = self.vr.dtype.names
            cols = {nm: np.ascontiguousarray(self.vr[nm]) for nm in names}
            rec_dtype = self.vr.dtype
        bx = bins(self.xmnsup, self.xsizsup, self.nxsup, cols['x'])
        by = bins(self.ymnsup, self.ysizsup, self.nysup, cols['y'])
        bz = bins(self.zmnsup, self.zsizsup, self.nzsup, cols['z'])
        blk = bx + by * self.nxsup + bz * (self.nxsup * self.nysup)
        self.block_of_sample = blk
        # stable: samples of one super block keep file order (the reference's default
        # argsort at super_block.py:135 leaves that order unspecified)
        if device is not None and torch.device(device).type == 'cuda' and nsup < 2 ** 31:
            key = torch.from_numpy(blk.astype(np.int32)).to(device)
            self.sort_index = torch.sort(key, stable=True)[1].cpu().numpy()
        else:
            self.sort_index = np.argsort(blk, kind='stable')
        # sorted columns (contiguous: what the device upload takes) and the sorted record
        # array the reference exposes (krige3d.py:690)
        self.cols = {nm: c[self.sort_index] for nm, c in cols.items()}
        self.vr = np.rec.fromarrays([self.cols[nm] for nm in names], dtype=rec_dtype)
        counts = np.bincount(blk, minlength=nsup)
        self.nisb = np.cumsum(counts, dtype=np.int64)
        self.sbstart = np.concatenate([[0], self.nisb]).astype(np.int32)

        def node_starts(mn, siz, n, mnsup, sizsup, nsup_axis):
            loc = mn + np.arange(n) * siz  # krige3d.py:307-309
            lut = np.searchsorted(_edges_node(mnsup, sizsup, nsup_axis), loc) - 1
            lut = np.clip(lut, 0, nsup_axis - 1)
            return lut.astype(np.int32), np.searchsorted(
                lut, np.arange(nsup_axis + 1), side='left').astype(np.int32)

        self.lutx, self.nodex0 = node_starts(self.xmn, self.xsiz, self.nx, self.xmnsup,
                                             self.xsizsup, self.nxsup)
        self.luty, self.nodey0 = node_starts(self.ymn, self.ysiz, self.ny, self.ymnsup,
                                             self.ysizsup, self.nysup)
        self.lutz, self.nodez0 = node_starts(self.zmn, self.zsiz, self.nz, self.zmnsup,
                                             self.zsizsup, self.nzsup)

    def block_of_locations(self, x, y, z):
        """getindx (super_block.py:319-346) for arbitrary locations: flat super-block index
        of each (x, y, z), clamped to the network (the reference does not clamp; results
        are only defined for locations inside the grid extent)."""
        out = 0
        mul = 1
        for c, mn, siz, n in ((x, self.xmnsup, self.xsizsup, self.nxsup),
                              (y, self.ymnsup, self.ysizsup, self.nysup),
                              (z, self.zmnsup, self.zsizsup, self.nzsup)):
            b = np.clip(bins_left(_edges_node(mn, siz, n), c), 0, n - 1)
            out = out + b * mul
            mul *= n
        return out

    def pickup(self, device=None):
        """Search template (super_block.py:140-161, func_pickup :226-265) evaluated by
        the k3d_build_template kernel, then folded into x-runs on the host."""
        device = torch.device(device if device is not None else "cuda")
        nx2, ny2, nz2 = 2 * self.nxsup - 1, 2 * self.nysup - 1, 2 * self.nzsup - 1
        mask = torch.empty(nx2 * ny2 * nz2, dtype=torch.uint8, device=device)
        rot = np.ascontiguousarray(np.asarray(self.rotmat, dtype=np.float64).ravel())
        with torch.cuda.device(device):
            stream = torch.cuda.current_stream().cuda_stream
            _native.check(_native.lib().k3d_build_template(
                self.nxsup, self.nysup, self.nzsup, float(self.xsizsup), float(self.ysizsup),
                float(self.zsizsup), rot.ctypes.data_as(C.POINTER(C.c_double)),
                float(self.radsqd), C.c_void_p(mask.data_ptr()), C.c_void_p(stream)))
        # the offsets that are set, in C order == the reference's x-outer, z-inner order
        idx = torch.nonzero(mask.view(nx2, ny2, nz2)).cpu().numpy()
        self.set_template_offsets(idx[:, 0] - (self.nxsup - 1), idx[:, 1] - (self.nysup - 1),
                                  idx[:, 2] - (self.nzsup - 1))

    def set_template_offsets(self, dx, dy, dz):
        """Template as offset lists (x outermost, z innermost, as func_pickup emits them)."""
        self.ixsbtosr, self.iysbtosr, self.izsbtosr = dx.tolist(), dy.tolist(), dz.tolist()
        self.nsbtosr = len(self.ixsbtosr)
        self.runs = runs_from_offsets(dx, dy, dz)

    def set_template_mask(self, m):
        """m[i, j, k] is True when offset (i-(nxsup-1), j-(nysup-1), k-(nzsup-1)) is searched."""
        i, j, k = np.nonzero(m)  # C order == the reference's x-outer, z-inner order
        self.ixsbtosr = (i - (self.nxsup - 1)).tolist()
        self.iysbtosr = (j - (self.nysup - 1)).tolist()
        self.izsbtosr = (k - (self.nzsup - 1)).tolist()
        self.nsbtosr = len(self.ixsbtosr)
        self.runs = template_runs(m, self.nxsup, self.nysup, self.nzsup)


def brick_template(dx, dy, dz, brick):
    """Template of a brick of brick[0] x brick[1] x brick[2] super blocks, relative to its
    first block: the union of the single-block template shifted to each block of the brick.
    Returns (runs, number of offsets)."""
    off = np.stack([np.asarray(a, dtype=np.int64) for a in (dx, dy, dz)], axis=1)
    i, j, k = np.meshgrid(np.arange(brick[0]), np.arange(brick[1]), np.arange(brick[2]),
                          indexing='ij')
    shifts = np.stack([i.ravel(), j.ravel(), k.ravel()], axis=1)
    allo = np.unique((off[:, None, :] + shifts[None, :, :]).reshape(-1, 3), axis=0)
    return runs_from_offsets(allo[:, 0], allo[:, 1], allo[:, 2]), int(allo.shape[0])


def runs_from_offsets(dx, dy, dz):
    """Same runs as template_runs, from the offset lists: sort by (dz, dy, dx) and cut
    wherever dx stops being consecutive."""
    dx, dy, dz = (np.asarray(a, dtype=np.int64) for a in (dx, dy, dz))
    if dx.size == 0:
        return np.zeros((0, 4), np.int16)
    o = np.lexsort((dx, dy, dz))
    x, y, z = dx[o], dy[o], dz[o]
    cut = np.ones(x.size, bool)
    cut[1:] = (z[1:] != z[:-1]) | (y[1:] != y[:-1]) | (x[1:] != x[:-1] + 1)
    first = np.nonzero(cut)[0]
    last = np.append(first[1:], x.size) - 1
    return np.ascontiguousarray(np.stack([x[first], x[last], y[first], z[first]],
                                         axis=1).astype(np.int16))


def template_runs(m, nxsup, nysup, nzsup):
    """Fold the boolean template m[i,j,k] into runs (dx0, dx1, dy, dz), ascending in
    (dz, dy, dx0): the union of the runs is exactly the template; each run is one
    contiguous range of flat super-block indices."""
    a = np.ascontiguousarray(np.transpose(m, (2, 1, 0))).astype(np.int8)  # [k, j, i]
    pad = np.zeros(a.shape[:2] + (1,), np.int8)
    d = np.diff(np.concatenate([pad, a, pad], axis=2), axis=2)
    ks, js, starts = np.nonzero(d == 1)
    ke, je, ends = np.nonzero(d == -1)
    assert np.array_equal(ks, ke) and np.array_equal(js, je)
    runs = np.stack([starts - (nxsup - 1), ends - 1 - (nxsup - 1), js - (nysup - 1),
                     ks - (nzsup - 1)], axis=1).astype(np.int16)
    return np.ascontiguousarray(runs.reshape(-1, 4))
