"""Krige3d -- drop-in for the reference's 3-D kriging program on a B200.

Same constructor, parameter files, attributes and driver method as
/root/reference/pygeostatistics/krige3d.py:27-411:

    k = Krige3d("testData/test_krige3d.par")
    k.kt3d()                      # or k.kriging()
    k.estimation, k.estimation_variance

What differs is where the node loop (krige3d.py:302-411) runs: the super-block search,
the kriging-system assembly and the solve are two sm_100a kernels reached through
the C ABI of include/k3d.h.  There is no CPU fallback.

Beyond the reference (it parses these but crashes on them):
universal kriging drift terms, kriging with external drift / non-stationary SK (`secfl`),
octant search (`noct`), block kriging (`nxdis/nydis/nzdis`), `itrend`, and the gslib
output file (`outfl`, via write_output()).
"""
import ctypes as C
import mmap
import os
import time
import warnings
import weakref
from collections import OrderedDict, namedtuple

import numpy as np
import torch

from . import _native
from .gslib_io import (SpatialData, read_params, read_grid_column, write_gslib_grid,
                       write_gslib_validation)
from .super_block import SuperBlockSearcher, brick_template

EPS = np.finfo(float).eps

_PINNED = []  # [pinned tensor, weakref to the ndarray it was lent to (or None)]
_PINNED_MAX = 8  # buffers kept (est, var, status and the three neighbour arrays of one run, + slack)


def release_host_buffers():
    """Drop the pooled page-locked result buffers that are not lent out, the shared result
    buffers of multi-rank runs nobody holds any more, the cached set-ups, and what the native
    library keeps between calls (k3d_release: scratch pool, timing events)."""
    _PINNED[:] = [e for e in _PINNED if not (e[1] is None or e[1]() is None)]
    for key in [k for k, sg in _SHARED.items() if sg.free()]:
        _SHARED.pop(key).close()
    clear_setup_cache()
    _native.check(_native.lib().k3d_release())



def _pinned_array(numel, dtype):
    """Page-locked host buffer for a result copy, as (tensor view, ndarray view).  Buffers
    are pooled: pinning (and first touching) hundreds of MB costs ~100 ms, several times
    the PCIe copy itself.  A buffer is lent out again only after the ndarray it backed has
    been garbage collected, so the arrays handed to the caller stay private."""
    entry = None
    for e in _PINNED:
        if e[0].dtype == dtype and e[0].numel() >= numel and (e[1] is None or e[1]() is None):
            entry = e
            break
    if entry is None:
        entry = [torch.empty(max(numel, 1), dtype=dtype, pin_memory=True), None]
        _PINNED.append(entry)
        if len(_PINNED) > _PINNED_MAX:  # drop idle buffers, oldest first
            idle = [id(e) for e in _PINNED[:-1] if e[1] is None or e[1]() is None]
            drop = set(idle[:len(_PINNED) - _PINNED_MAX])
            _PINNED[:] = [e for e in _PINNED if id(e) not in drop]  # (list.remove would compare tensors)
    view = entry[0][:numel]
    arr = view.numpy()
    entry[1] = weakref.ref(arr)
    return view, arr


def slab_ranges(nz, world_size):
    """Contiguous z-plane ranges [iz0, iz1), one per rank, sizes differing by at most 1."""
    base, extra = divmod(nz, world_size)
    out, z = [], 0
    for r in range(world_size):
        n = base + (1 if r < extra else 0)
        out.append((z, z + n))
        z += n
    return out


def gather_slabs(local, ranges, plane, group=None):
    """All-gather per-rank slab arrays (1-D tensors, rank r holding planes ranges[r]) into
    the full grid on every rank.  Slabs are padded to the largest slab so a single
    all_gather_into_tensor moves everything (NCCL over NVLink on GPUs, gloo on CPU)."""
    import torch.distributed as dist
    world = len(ranges)
    maxp = max(b - a for a, b in ranges)
    if all(b - a == maxp for a, b in ranges):  # equal slabs: straight into the result
        full = torch.empty(world * maxp * plane, dtype=local.dtype, device=local.device)
        dist.all_gather_into_tensor(full, local.contiguous(), group=group)
        return full
    pad = torch.zeros(maxp * plane, dtype=local.dtype, device=local.device)
    pad[:local.numel()] = local
    full = torch.empty(world * maxp * plane, dtype=local.dtype, device=local.device)
    dist.all_gather_into_tensor(full, pad, group=group)
    parts = [full[r * maxp * plane: r * maxp * plane + (b - a) * plane]
             for r, (a, b) in enumerate(ranges)]
    return torch.cat(parts)


_FP_WEIGHTS = {}


def data_fingerprint(vr):
    """Cheap checksum of the columns of a record array (or of a mapping name -> column): per
    column the sum of its 64-bit words and their dot product with odd position-dependent
    weights, modulo 2^64 -- any edit of a value and any re-ordering of the rows changes it."""
    names = list(vr.dtype.names) if hasattr(vr, 'dtype') else list(vr.keys())
    out = []
    for nm in names:
        words = np.ascontiguousarray(vr[nm], dtype=np.float64).view(np.uint64)
        w = _FP_WEIGHTS.get(words.size)
        if w is None:
            _FP_WEIGHTS.clear()  # (one size at a time)
            w = _FP_WEIGHTS[words.size] = (np.arange(words.size, dtype=np.uint64) * np.uint64(2) +
                                           np.uint64(1)) * np.uint64(0x9E3779B97F4A7C15)
        out.append((nm, int(words.size), int(words.sum(dtype=np.uint64)), int(np.dot(words, w))))
    return tuple(out)


_SEARCHERS = OrderedDict()  # set-up key -> SuperBlockSearcher (binning, sort, template, pinned inputs)
_SEARCHER_CACHE_SIZE = 4


def clear_setup_cache():
    """Forget the cached super-block set-ups (binning, sort, search template, pinned copies)."""
    _SEARCHERS.clear()


class _SharedGrid(object):
    """Result arrays of a multi-rank run in ONE host buffer shared by the ranks of the node
    (a file in /dev/shm mapped by every rank and unlinked as soon as all have it): each rank
    copies only its own z-slab from its GPU, and every rank reads the whole grid.  The part a
    rank writes is page-locked (cudaHostRegister), so the copies are asynchronous DMA."""

    def __init__(self, spec, rank, group):
        import torch.distributed as dist
        self.layout, off = {}, 0
        for name, dt, n in spec:  # every array starts on a page
            self.layout[name] = (np.dtype(dt), n, off)
            off += -(-np.dtype(dt).itemsize * max(n, 1) // 4096) * 4096
        self.nbytes = off
        path = [None]
        if rank == 0:
            path[0] = "/dev/shm/k3d_%d_%x" % (os.getpid(), int(time.time() * 1e6) & 0xffffffff)
            with open(path[0], "wb") as f:
                f.truncate(self.nbytes)
        src = dist.get_global_rank(group, 0) if group is not None else 0
        dist.broadcast_object_list(path, src=src, group=group)
        ok, self.map = 1, None
        try:
            with open(path[0], "r+b") as f:
                self.map = mmap.mmap(f.fileno(), self.nbytes)
        except (OSError, ValueError) as e:
            ok = 0
            if os.environ.get("K3D_DEBUG"):
                print("k3d: rank %d cannot map %r: %r" % (rank, path[0], e), flush=True)
        flag = torch.tensor([ok], dtype=torch.int32,
                            device="cuda" if dist.get_backend(group) == "nccl" else "cpu")
        dist.all_reduce(flag, op=dist.ReduceOp.MIN, group=group)
        # reading the result is what makes the host wait (NCCL only queues the reduction):
        # after it every rank has mapped the file, and the name can go
        self.ok = bool(int(flag.item()))
        if rank == 0:
            os.unlink(path[0])
        if os.environ.get("K3D_DEBUG"):
            print("k3d: rank %d shared grid %r local ok %d, all ok %d" % (rank, path[0], ok, self.ok), flush=True)
        self.registered = []
        self.lent = None  # weakref to the array the last result was a view of
        self.addr = np.frombuffer(self.map, dtype=np.uint8).ctypes.data if self.map is not None else 0

    def free(self):
        return self.lent is None or self.lent() is None

    def lend(self):
        """A fresh array over the buffer: it dies with the last result view cut from it."""
        base = np.frombuffer(self.map, dtype=np.uint8)
        self.lent = weakref.ref(base)
        return base

    def view(self, base, name):
        dt, n, off = self.layout[name]
        return base[off: off + dt.itemsize * n].view(dt)

    def pin(self, name, a, b):
        """Page-lock elements [a, b) of array `name` (whole pages around them), once."""
        dt, n, off = self.layout[name]
        lo = (off + a * dt.itemsize) // 4096 * 4096
        hi = min(self.nbytes, -(-(off + b * dt.itemsize) // 4096) * 4096)
        if hi <= lo or not torch.cuda.is_available():
            return
        for rlo, rhi in self.registered:
            if rlo <= lo and hi <= rhi:
                return
        err = torch.cuda.cudart().cudaHostRegister(self.addr + lo, hi - lo, 0)
        if int(err) != 0:
            raise _native.K3dError("cudaHostRegister failed: %s" % err)
        self.registered.append((lo, hi))

    def close(self):
        if torch.cuda.is_available():
            for lo, hi in self.registered:
                torch.cuda.cudart().cudaHostUnregister(self.addr + lo)
        self.registered = []


_SHARED = {}  # (world, rank, layout key) -> _SharedGrid


class _MemoryData(object):
    """SpatialData look-alike over in-memory columns (same record-array layout, built when
    somebody asks for it; the set-up works from the columns)."""

    def __init__(self, columns):
        names = list(columns.keys())
        arrays = [np.ascontiguousarray(columns[n], dtype=np.float64) for n in names]
        self.datafl = None
        self.property_name = [n for n in names if n not in ('x', 'y', 'z')]
        self._2d = 'z' not in names
        if self._2d:
            names.append('z')
            arrays.append(np.zeros_like(arrays[0]))
        self.columns = OrderedDict(zip(names, arrays))
        self._vr = None

    @property
    def vr(self):
        if self._vr is None:
            names = list(self.columns.keys())
            self._vr = np.rec.fromarrays(list(self.columns.values()), dtype=np.dtype(
                {'names': names, 'formats': ['f8'] * len(names)}))
        return self._vr


class Krige3d(object):
    """Performing 3d Kriging with super block search (GPU)."""

    def __init__(self, param_file, data=None, secondary=None, jackknife=None):
        """`param_file`: path of a JSON .par file (as in the reference), or the same
        mapping already parsed.  `data` (optional): ordered mapping column name -> 1-D
        array used instead of reading `datafl`; `secondary` (optional): the external-drift
        value of every grid node (x fastest) used instead of reading `secfl`; `jackknife`
        (optional, `option` 2): the same kind of mapping used instead of reading `jackfl`."""
        self.param_file = param_file
        self._read_params()
        self._check_params()
        self.data = SpatialData(self.datafl) if data is None else _MemoryData(data)
        self._secondary = secondary
        self._jackknife = jackknife
        self.locations = None      # option 1/2: (n, 3) estimated locations, file order
        self.true_values = None    # option 1/2: the value observed at each location
        self.property_name = self.data.property_name
        self._vr = None
        self.rotmat = None
        self.estimation = None
        self.estimation_variance = None
        self.status = None
        self.neighbours = None
        self.neighbour_count = None
        self.xdb = self.ydb = self.zdb = None
        self._2d = self.data._2d
        self.searcher = None
        self.const = None
        self._block_covariance = None
        self.unbias = None
        self.maxcov = None
        self._mdt = None
        self.resc = None
        self.verbose = True
        self.timings = {}
        self._dev = None
        self.brick_size = None         # (bx, by, bz) super blocks per CTA tile; None = automatic
        self.search_identity = False   # Krige2d: isotropic search with the exact identity matrix
        self.flags = 0                 # K3dParams.flags (include/k3d.h)

    # ------------------------------------------------------------------ parameters
    def _read_params(self):
        params = self.param_file if isinstance(self.param_file, dict) \
            else read_params(self.param_file)
        g = params.__getitem__  # KeyError for a missing key, like the reference
        self.datafl = g('datafl')
        self.ixl, self.iyl, self.izl = g('icolx'), g('icoly'), g('icolz')
        self.ivrl, self.iextv = g('icolvr'), g('icolsec')
        self.tmin, self.tmax = g('tmin'), g('tmax')
        self.koption = g('option')
        self.jackfl = g('jackfl')
        self.ixlj, self.iylj, self.izlj = g('jicolx'), g('jicoly'), g('jicolz')
        self.ivrlj, self.iextvj = g('jicolvr'), g('jicolsec')
        self.idbg, self.dbgfl, self.outfl = g('idbg'), g('dbgfl'), g('outfl')
        self.nx, self.xmn, self.xsiz = g('nx'), g('xmn'), g('xsiz')
        self.ny, self.ymn, self.ysiz = g('ny'), g('ymn'), g('ysiz')
        self.nz, self.zmn, self.zsiz = g('nz'), g('zmn'), g('zsiz')
        self.nxdis, self.nydis, self.nzdis = g('nxdis'), g('nydis'), g('nzdis')
        self.ndmin, self.ndmax, self.noct = g('ndmin'), g('ndmax'), g('noct')
        self.radius_hmax, self.radius_hmin = g('radius_hmax'), g('radius_hmin')
        self.radius_vert = g('radius_vert')
        self.sang1, self.sang2, self.sang3 = g('sang1'), g('sang2'), g('sang3')
        self.ktype, self.skmean = g('ikrige'), g('skmean')
        self.idrift, self.itrend = list(g('idrift')), g('itrend')
        self.extfl, self.iextve = g('secfl'), g('iseccol')
        self.nst, self.c0 = g('nst'), g('c0')
        self.it, self.cc = list(g('it')), list(g('cc'))
        self.ang1, self.ang2, self.ang3 = list(g('ang1')), list(g('ang2')), list(g('ang3'))
        self.aa_hmax, self.aa_hmin = list(g('aa_hmax')), list(g('aa_hmin'))
        self.aa_vert = list(g('aa_vert'))

    def _check_params(self):
        # krige3d.py:132-159 (same messages)
        if self.radius_hmax <= 0:
            raise ValueError("radius_hmax should be larger than zero.")
        if self.nst <= 0:
            raise ValueError("nst must be at least 1.")
        for vtype, a_range in zip(self.it, self.aa_hmax):
            if vtype not in (1, 2, 3, 4, 5):
                raise ValueError("INVALID variogram number {}".format(vtype))
            if vtype == 4 and (a_range < 0 or a_range > 2.0):
                raise ValueError("INVALID power variogram")
            if vtype != 4 and not a_range > 0:  # (the reference divides by it: inf / nan everywhere)
                raise ValueError("variogram range aa_hmax must be positive")
        if self.ktype == 3 and self.iextv <= 0:
            raise ValueError("Must have exteranl variable")
        if self.ixl < 0 and self.nx > 1:
            raise ValueError("WARNING: ixl=0 and nx>1 !")
        if self.iyl < 0 and self.ny > 1:
            raise ValueError("WARNING: iyl=0 and ny>1 !")
        if self.izl < 0 and self.nz > 1:
            raise ValueError("WARNING: izl=0 and nz>1 !")
        for item in self.idrift:
            if not isinstance(item, bool):
                raise ValueError("Invalid drift term {}".format(item))
        # limits of this build (include/k3d.h)
        if self.nst > _native.MAXNST:
            raise ValueError("nst > {} is not supported".format(_native.MAXNST))
        if not 1 <= self.ndmax <= _native.MAXNDMAX:
            raise ValueError("ndmax must be in 1..{}".format(_native.MAXNDMAX))
        if self.koption not in (0, 1, 2):
            raise ValueError("option must be 0 (grid), 1 (cross-validation) or 2 (jackknife)")

    def read_data(self):
        """(Re)load `datafl`.  The reference does this in the constructor."""
        self.data = SpatialData(self.datafl)
        self.property_name = self.data.property_name
        self._vr = None
        self._2d = self.data._2d

    @property
    def vr(self):
        """The data records: in file order until the run is prepared, then sorted by super
        block (krige3d.py:690)."""
        return self.data.vr if self._vr is None else self._vr

    @vr.setter
    def vr(self, value):
        self._vr = value

    # ------------------------------------------------------------------ once-per-run set-up
    def _preprocess(self):
        const = namedtuple('Krige3d_const', ['PMX', 'MAXNST', 'MAXDT', 'MAXSB', 'MAXDIS',
                                             'MAXSAM', 'UNEST'])

        def maxsb(n):
            return min(int(n / 2), 50) if n > 1 else 1

        self.const = const(PMX=999, MAXNST=4, MAXDT=9,
                           MAXSB=(maxsb(self.nx), maxsb(self.ny), maxsb(self.nz)),
                           MAXDIS=self.nxdis * self.nydis * self.nzdis,
                           MAXSAM=self.ndmax + 1, UNEST=np.nan)
        self.radsqd = self.radius_hmax * self.radius_hmax
        self.sanis1 = self.radius_hmin / self.radius_hmax
        self.sanis2 = self.radius_vert / self.radius_hmax
        self.anis1 = np.array(self.aa_hmin) / np.maximum(self.aa_hmax, EPS)
        self.anis2 = np.array(self.aa_vert) / np.maximum(self.aa_hmax, EPS)
        if self.koption == 1:  # cross-validation: the data file is the jackknife file
            self.jackfl = self.datafl  # krige3d.py:200-208
            self.ixlj, self.iylj, self.izlj = self.ixl, self.iyl, self.izl
            self.ivrlj, self.iextvj = self.ivrl, self.iextv

    def _set_rotation(self):
        """rotmat[(nst+1),3,3]; the last matrix is the search ellipsoid (krige3d.py:211-253).
        The azimuth test `ang <= 0 and ang < 270` is the reference's, kept so that the
        matrices -- and hence every search distance -- are bit-identical."""
        ang1 = np.append(self.ang1, self.sang1)
        ang2 = np.append(self.ang2, self.sang2)
        ang3 = np.append(self.ang3, self.sang3)
        anis1 = np.append(self.anis1, self.sanis1)
        anis2 = np.append(self.anis2, self.sanis2)
        alpha = np.array([np.deg2rad(90 - a) if (a <= 0 and a < 270) else np.deg2rad(450 - a)
                          for a in ang1])
        beta, theta = np.deg2rad(-ang2), np.deg2rad(ang3)
        sina, sinb, sint = np.sin(alpha), np.sin(beta), np.sin(theta)
        cosa, cosb, cost = np.cos(alpha), np.cos(beta), np.cos(theta)
        afac1 = 1.0 / np.maximum(anis1, EPS)
        afac2 = 1.0 / np.maximum(anis2, EPS)
        rot = np.full((ang1.shape[0], 3, 3), np.nan)
        rot[:, 0, 0] = cosb * cosa
        rot[:, 0, 1] = cosb * sina
        rot[:, 0, 2] = -sinb
        rot[:, 1, 0] = afac1 * (-cost * sina + sint * sinb * cosa)
        rot[:, 1, 1] = afac1 * (cost * cosa + sint * sinb * sina)
        rot[:, 1, 2] = afac1 * (sint * cosb)
        rot[:, 2, 0] = afac2 * (sint * sina + cost * sinb * cosa)
        rot[:, 2, 1] = afac2 * (-sint * cosa + cost * sinb * sina)
        rot[:, 2, 2] = afac2 * (cost * cosb)
        if self.search_identity:  # krige2d front end: plain dx*dx + dy*dy (krige2d.py:318-320)
            rot[-1] = np.eye(3)
        self.rotmat = rot

    def _max_covariance(self):
        self.maxcov = self.c0
        for ist in range(self.nst):
            self.maxcov += self.const.PMX if self.it[ist] == 4 else self.cc[ist]

    def _rescaling(self):
        if self.radsqd < 1:
            self.resc = 2 * self.radius_hmax / max(self.maxcov, 0.0001)
        else:
            self.resc = (4 * self.radsqd) / max(self.maxcov, 0.0001)
        if self.resc <= 0:
            raise ValueError("rescaling value is wrong, {}".format(self.resc))
        self.resc = 1 / self.resc

    @property
    def mdt(self):
        """Number of constraint rows (krige3d.py:636-662)."""
        if self._mdt is None:
            if self.ktype == 0 or self.ktype == 2:
                self.idrift = [False] * 9
                self._mdt = 0
            else:
                self._mdt = 1 + sum(1 for b in self.idrift if b) + (1 if self.ktype == 3 else 0)
        return self._mdt

    def _create_searcher(self):
        """Super-block index of the data (krige3d.py:664-690).  Binning, the sort, the search
        template and the pinned copies the upload reads depend only on the grid, the search
        ellipsoid and the data: they are kept (a few set-ups, keyed by those parameters and a
        checksum of the data bytes), so repeated runs pay for them once."""
        # always from the records in file order
        cols = getattr(self.data, 'columns', None)
        if cols is None:
            cols = OrderedDict((nm, np.ascontiguousarray(self.data.vr[nm])) for nm in self.data.vr.dtype.names)
        dev = self._device()
        key = (self.nx, self.xmn, self.xsiz, self.ny, self.ymn, self.ysiz, self.nz, self.zmn,
               self.zsiz, tuple(self.const.MAXSB), self.rotmat[-1].tobytes(), self.radsqd,
               self.noct, str(dev), data_fingerprint(cols))
        s = _SEARCHERS.get(key)
        if s is None:
            s = SuperBlockSearcher()
            s.nx, s.xmn, s.xsiz = self.nx, self.xmn, self.xsiz
            s.ny, s.ymn, s.ysiz = self.ny, self.ymn, self.ysiz
            s.nz, s.zmn, s.zsiz = self.nz, self.zmn, self.zsiz
            s.columns = cols
            s.MAXSB = self.const.MAXSB
            s.rotmat = self.rotmat[-1]
            s.radsqd = self.radsqd
            s.noct = self.noct
            s.setup(dev)
            s.pickup(dev)
            s.pinned = {}
            _SEARCHERS[key] = s
            while len(_SEARCHERS) > _SEARCHER_CACHE_SIZE:
                _SEARCHERS.popitem(last=False)
        else:
            _SEARCHERS.move_to_end(key)
        self.searcher = s
        self.vr = s.vr  # sorted by super block (krige3d.py:690)

    def _block_discretization(self):
        """ndb discretisation points of a block, offsets from its centre
        (krige3d.py:692-707; the three axes are combined as a tensor product, x outer)."""
        self.nxdis = max(1, self.nxdis)
        self.nydis = max(1, self.nydis)
        self.nzdis = max(1, self.nzdis)
        self.ndb = self.nxdis * self.nydis * self.nzdis
        if self.ndb > _native.MAXDIS:
            raise ValueError("Too many discretization points")
        xdis, ydis, zdis = self.xsiz / self.nxdis, self.ysiz / self.nydis, self.zsiz / self.nzdis
        ax = np.arange(0, self.nxdis, 1) * xdis + (-0.5 * self.xsiz + 0.5 * xdis)
        ay = np.arange(0, self.nydis, 1) * ydis + (-0.5 * self.ysiz + 0.5 * ydis)
        az = np.arange(0, self.nzdis, 1) * zdis + (-0.5 * self.zsiz + 0.5 * zdis)
        gx, gy, gz = np.meshgrid(ax, ay, az, indexing='ij')
        self.xdb, self.ydb, self.zdb = gx.ravel(), gy.ravel(), gz.ravel()

    def _cova3_host(self, p1, p2):
        """cova3 (krige3d.py:838-875) in numpy -- only for the once-per-run block covariance."""
        d = np.asarray(p1, float) - np.asarray(p2, float)

        def sq(r):
            c = r @ d
            return float(c @ c)
        hsqd = sq(self.rotmat[0])
        if hsqd < EPS:
            return self.maxcov
        cova = 0.0
        for i in range(self.nst):
            h = np.sqrt(sq(self.rotmat[i]))
            a, c, t = self.aa_hmax[i], self.cc[i], self.it[i]
            if t == 1:
                hr = h / a
                if hr < 1:
                    cova += c * (1 - hr * (1.5 - 0.5 * hr * hr))
            elif t == 2:
                cova += c * np.exp(-3.0 * h / a)
            elif t == 3:
                cova += c * np.exp(-3.0 * (h / a) * (h / a))
            elif t == 4:
                cova += self.maxcov - c * (h ** a)
            elif t == 5:
                cova += c * np.cos(h / a * np.pi)
        return cova

    @property
    def block_covariance(self):
        """Average covariance within the block (krige3d.py:616-634)."""
        if self._block_covariance is None:
            if self.ndb <= 1:
                self._block_covariance = self.unbias
            else:
                pts = np.stack([self.xdb, self.ydb, self.zdb], axis=1)
                cov = np.array([[self._cova3_host(a, b) for b in pts] for a in pts])
                cov[np.diag_indices_from(cov)] -= self.c0
                self._block_covariance = float(np.mean(cov))
        return self._block_covariance

    # ------------------------------------------------------------------ device plumbing
    def _device(self):
        if self._dev is None:
            if not torch.cuda.is_available():
                raise _native.K3dError("Krige3d needs a CUDA device (B200); there is no CPU path")
            self._dev = torch.device("cuda", torch.cuda.current_device())
        return self._dev

    def _params_struct(self):
        P = _native.K3dParams()
        P.nx, P.ny, P.nz = self.nx, self.ny, self.nz
        P.xmn, P.ymn, P.zmn = self.xmn, self.ymn, self.zmn
        P.xsiz, P.ysiz, P.zsiz = self.xsiz, self.ysiz, self.zsiz
        P.ndmin, P.ndmax, P.noct = self.ndmin, self.ndmax, self.noct
        P.flags = int(self.flags)
        P.radsqd = self.radsqd
        P.ktype, P.itrend = self.ktype, int(bool(self.itrend))
        P.mdt = self.mdt
        for i in range(9):
            P.idrift[i] = int(bool(self.idrift[i]))
        P.skmean, P.tmin, P.tmax = self.skmean, self.tmin, self.tmax
        P.nst, P.c0 = self.nst, self.c0
        for i in range(self.nst):
            P.it[i], P.cc[i], P.aa[i] = int(self.it[i]), self.cc[i], self.aa_hmax[i]
        for i, v in enumerate(self.rotmat.ravel()):
            P.rotmat[i] = v
        P.maxcov, P.resc, P.unbias, P.cbb = self.maxcov, self.resc, self.unbias, \
            self.block_covariance
        for i in range(9):
            P.bv[i] = self.bv[i]
        s = self.searcher
        P.nxsup, P.nysup, P.nzsup, P.ndb = s.nxsup, s.nysup, s.nzsup, self.ndb
        for i, b in enumerate(self._brick()):
            P.brick[i] = b
        return P

    def _brick(self):
        """Super blocks per CTA tile and axis.  The reference sizes super blocks by the grid
        (at most 50 per axis, krige3d.py:164-170), so a small grid gets super blocks of a
        few nodes; CTAs then work on bricks of 2-4 blocks per axis (about 64 nodes or
        more), searching the union of their templates.  Grid runs only."""
        s = self.searcher
        if self.koption != 0:
            return (1, 1, 1)
        if self.brick_size is not None:
            return tuple(int(min(max(v, 1), n, 8)) for v, n in
                         zip(self.brick_size, (s.nxsup, s.nysup, s.nzsup)))
        tile = self.nx * self.ny * self.nz / float(s.nxsup * s.nysup * s.nzsup)
        if tile >= 48:
            return (1, 1, 1)
        for b in (2, 3, 4):
            dims = tuple(min(b, n) for n in (s.nxsup, s.nysup, s.nzsup))
            if tile * dims[0] * dims[1] * dims[2] >= 64:
                break
        return dims

    def _host_inputs(self):
        """Pinned host copies of everything the kernel reads (those of the sample set and its
        index live with the cached set-up)."""
        s = self.searcher
        c = s.cols  # contiguous sorted columns
        names = dict(sx='x', sy='y', sz='z', sv=self.property_name[0])
        if self.ktype in (2, 3):
            names['ssec'] = self.property_name[1]
        host = {}
        for k, nm in names.items():
            if ('col', nm) not in s.pinned:
                s.pinned[('col', nm)] = torch.from_numpy(c[nm]).pin_memory()
            host[k] = s.pinned[('col', nm)]
        host.setdefault('ssec', None)
        for k in ('sbstart', 'nodex0', 'nodey0', 'nodez0'):
            if k not in s.pinned:
                s.pinned[k] = torch.from_numpy(np.ascontiguousarray(getattr(s, k)).reshape(-1)).pin_memory()
            host[k] = s.pinned[k]
        brick = self._brick()
        if ('runs', brick) not in s.pinned:  # the template of a tile (a brick of super blocks)
            runs, ntmpl = (s.runs, s.nsbtosr) if brick == (1, 1, 1) else \
                brick_template(s.ixsbtosr, s.iysbtosr, s.izsbtosr, brick)
            s.pinned[('runs', brick)] = (torch.from_numpy(np.ascontiguousarray(runs).reshape(-1)).pin_memory(),
                                         int(ntmpl))
        host['runs'], self._ntmpl = s.pinned[('runs', brick)]
        host['dbxyz'] = torch.from_numpy(
            np.ascontiguousarray(np.concatenate([self.xdb, self.ydb, self.zdb]))).pin_memory()
        return host

    def _upload(self, host, ext_slab):
        dev = self._device()
        d = {k: (None if v is None else v.to(dev, non_blocking=True)) for k, v in host.items()}
        d['extgrid'] = None if ext_slab is None else ext_slab.to(dev, non_blocking=True)
        return d

    def _inputs_struct(self, d):
        ptr = lambda t: C.c_void_p(None if t is None else t.data_ptr())
        I = _native.K3dInputs()
        I.nd = int(d['sx'].numel())
        I.sx, I.sy, I.sz, I.sv, I.ssec = ptr(d['sx']), ptr(d['sy']), ptr(d['sz']), ptr(d['sv']), \
            ptr(d['ssec'])
        I.sbstart, I.nruns, I.runs = ptr(d['sbstart']), int(d['runs'].numel() // 4), ptr(d['runs'])
        I.ntmpl = int(getattr(self, '_ntmpl', None) or self.searcher.nsbtosr)
        I.nodex0, I.nodey0, I.nodez0 = ptr(d['nodex0']), ptr(d['nodey0']), ptr(d['nodez0'])
        I.dbxyz, I.extgrid = ptr(d['dbxyz']), ptr(d['extgrid'])
        return I

    def _outputs_struct(self, out):
        ptr = lambda t: C.c_void_p(None if t is None else t.data_ptr())
        O = _native.K3dOutputs()
        O.est, O.var = ptr(out['est']), ptr(out['var'])
        O.nbr_idx, O.nbr_cnt, O.status = ptr(out.get('nbr_idx')), ptr(out.get('nbr_cnt')), \
            ptr(out.get('status'))
        O.ncand = ptr(out.get('ncand'))
        return O

    def _launch(self, P, d, iz0, iz1, out):
        I, O = self._inputs_struct(d), self._outputs_struct(out)
        stream = torch.cuda.current_stream(self._device()).cuda_stream
        _native.check(_native.lib().k3d_krige_slab(C.byref(P), C.byref(I), iz0, iz1,
                                                  C.byref(O), C.c_void_p(stream)))

    def _alloc_outputs(self, nodes, neighbours):
        dev = self._device()
        out = dict(est=torch.empty(nodes, dtype=torch.float64, device=dev),
                   var=torch.empty(nodes, dtype=torch.float64, device=dev),
                   status=torch.empty(nodes, dtype=torch.uint8, device=dev))
        if neighbours:
            out['nbr_idx'] = torch.empty(nodes * self.ndmax, dtype=torch.int32, device=dev)
            out['nbr_cnt'] = torch.empty(nodes, dtype=torch.int32, device=dev)
            out['ncand'] = torch.empty(nodes, dtype=torch.int32, device=dev)
        return out

    def _external_grid(self):
        if self.ktype not in (2, 3):
            return None
        ext = np.ascontiguousarray(self._secondary, dtype=np.float64) \
            if self._secondary is not None else read_grid_column(self.extfl, self.iextve)
        if ext.shape[0] != self.nx * self.ny * self.nz:
            raise ValueError("secfl holds {} values, grid has {} nodes".format(
                ext.shape[0], self.nx * self.ny * self.nz))
        return torch.from_numpy(np.array(ext, dtype=np.float64, copy=True))

    _COPY_STREAMS = {}

    def _copy_stream(self):
        dev = self._device()
        if dev not in Krige3d._COPY_STREAMS:
            Krige3d._COPY_STREAMS[dev] = torch.cuda.Stream(dev)
        return Krige3d._COPY_STREAMS[dev]

    def _plane_runs(self, iz0, iz1, target=None):
        """[iz0, iz1) cut into about `target` runs of z-planes at super-block boundaries (a
        run that splits a super block would shorten the walks of its tiles)."""
        if target is None:
            target = int(os.environ.get("K3D_PLANE_RUNS", "4"))
        starts = [int(z) for z in self.searcher.nodez0 if iz0 < z < iz1]
        want = [iz0 + (iz1 - iz0) * j // target for j in range(1, target)]
        cuts = sorted({min(starts, key=lambda z: abs(z - w)) for w in want}) if starts else []
        edges = [iz0] + cuts + [iz1]
        return [(a, b) for a, b in zip(edges[:-1], edges[1:]) if b > a]

    # ------------------------------------------------------------------ the driver
    def prepare(self):
        """Everything krige3d.py:255-282 does before the node loop."""
        t0 = time.time()
        self._preprocess()
        self._set_rotation()
        self._max_covariance()
        self._rescaling()
        _ = self.mdt
        if self.verbose:
            print("Setting up Super Block Search...")
        self._create_searcher()
        self._block_discretization()
        self.unbias = self.maxcov
        self.bv = np.array([
            np.mean(self.xdb), np.mean(self.ydb), np.mean(self.zdb),
            np.mean(self.xdb * self.xdb), np.mean(self.ydb * self.ydb),
            np.mean(self.zdb * self.zdb), np.mean(self.xdb * self.ydb),
            np.mean(self.xdb * self.zdb), np.mean(self.ydb * self.zdb)]) * self.resc
        self.timings['setup_s'] = time.time() - t0

    def kt3d(self, neighbours=False, group=None, gather="host"):
        """Krige the whole grid.  Results: `estimation`, `estimation_variance` (numpy
        float64, length nx*ny*nz, x fastest, NaN = unestimated) and `status` (uint8 codes
        of include/k3d.h).  With neighbours=True also `neighbours` [nodes, ndmax] (original
        sample row numbers, nearest first, -1 padded) and `neighbour_count`.

        Under an initialised torch.distributed process group the grid is split into
        z-slabs, one per rank, and every rank ends up with the full arrays.  gather="host"
        (default): each rank copies only its own slab from its GPU, into one host buffer the
        ranks of the node share (falls back to "device" when they do not share /dev/shm);
        gather="device": the slabs are all-gathered between the GPUs (NCCL) and every rank
        copies the whole grid to its host."""
        import torch.distributed as dist
        self.prepare()
        if self.koption != 0:
            return self._kt3d_locations(neighbours)
        P = self._params_struct()
        world, rank = 1, 0
        if dist.is_available() and dist.is_initialized():
            world, rank = dist.get_world_size(group), dist.get_rank(group)
        ranges = slab_ranges(self.nz, world)
        iz0, iz1 = ranges[rank]
        plane = self.nx * self.ny
        nodes = (iz1 - iz0) * plane
        if self.verbose:
            print("Start working on the kriging...")
        t0 = time.time()
        host = self._host_inputs()
        ext = self._external_grid()
        ext_slab = None if ext is None else ext[iz0 * plane: iz1 * plane].pin_memory()
        d = self._upload(host, ext_slab)
        out = self._alloc_outputs(nodes, neighbours)
        width = lambda k: self.ndmax if k == 'nbr_idx' else 1
        shared = None
        tm = self.timings
        tm['upload_s'] = time.time() - t0
        if world > 1 and gather == "host":
            shared = self._shared_grid(out, self.nz * plane, world, rank, group)
        tm['shared_s'] = time.time() - t0 - tm['upload_s']
        res = {}
        if world == 1 or shared is not None:
            # The slab goes in a few runs of z-planes (cut where super blocks end), and the
            # device-to-host copy of a run overlaps the kernels of the next.  One rank: into
            # pooled pinned buffers; several: into the slab's place in the shared buffer.
            dev = self._device()
            main, side = torch.cuda.current_stream(dev), self._copy_stream()
            if shared is None:
                pinned = {k: _pinned_array(v.numel(), v.dtype) for k, v in out.items()}
                dst = {k: pv[0] for k, pv in pinned.items()}
                res = {k: pv[1] for k, pv in pinned.items()}
            else:
                base = shared.lend()
                res = {k: shared.view(base, k) for k in out}
                dst = {}
                for k in out:
                    a, b = iz0 * plane * width(k), iz1 * plane * width(k)
                    shared.pin(k, a, b)
                    dst[k] = torch.from_numpy(res[k][a:b])
            for za, zb in (self._plane_runs(iz0, iz1) if nodes > 0 else []):
                a, b = (za - iz0) * plane, (zb - iz0) * plane
                dd = dict(d)
                if d['extgrid'] is not None:
                    dd['extgrid'] = d['extgrid'][a:b]
                self._launch(P, dd, za, zb, {k: v[a * width(k): b * width(k)] for k, v in out.items()})
                done = torch.cuda.Event()
                done.record(main)
                side.wait_event(done)
                with torch.cuda.stream(side):
                    for k, v in out.items():
                        dst[k][a * width(k): b * width(k)].copy_(
                            v[a * width(k): b * width(k)], non_blocking=True)
            tm['launch_s'] = time.time() - t0 - tm['upload_s'] - tm['shared_s']
            side.synchronize()
            main.wait_stream(side)
            tm['device_wait_s'] = time.time() - t0 - tm['upload_s'] - tm['shared_s'] - tm['launch_s']
            if shared is not None:
                t1 = time.time()
                dist.barrier(group)  # every rank's slab is in place
                tm['barrier_s'] = time.time() - t1
        else:
            if nodes > 0:
                self._launch(P, d, iz0, iz1, out)
            out = {k: gather_slabs(v, ranges, plane * width(k), group) for k, v in out.items()}
            for k, v in out.items():
                view, arr = _pinned_array(v.numel(), v.dtype)
                view.copy_(v, non_blocking=True)
                res[k] = arr
        torch.cuda.synchronize(self._device())
        self.timings['krige_s'] = time.time() - t0
        self.timings['total_s'] = time.time() - t0 + self.timings.get('setup_s', 0.0)
        self.estimation = res['est']
        self.estimation_variance = res['var']
        self.status = res['status']
        if neighbours:
            idx = res['nbr_idx'].reshape(-1, self.ndmax).astype(np.int64)
            self.neighbour_count = res['nbr_cnt']
            self.candidates_tested = res['ncand']
            orig = np.where(idx >= 0, self.searcher.sort_index[np.maximum(idx, 0)], -1)
            self.neighbours = orig
            self.neighbours_sorted_space = idx
        if self.verbose:
            print("Kriging Finished.")

    def _shared_grid(self, out, total, world, rank, group):
        """The host buffer the ranks share for results of this shape, or None when they cannot
        share one.  A buffer is used again only when EVERY rank has let go of the arrays the
        previous run handed out (a collective decision, so the ranks stay in step)."""
        import torch.distributed as dist
        spec = tuple((k, str(v.dtype).replace('torch.', ''), total * (self.ndmax if k == 'nbr_idx' else 1))
                     for k, v in out.items())
        key = (world, rank, id(group), spec)
        sg = _SHARED.get(key)
        nccl = dist.get_backend(group) == "nccl"
        flag = torch.tensor([1 if (sg is not None and sg.free()) else 0], dtype=torch.int32,
                            device=self._device() if nccl else "cpu")
        dist.all_reduce(flag, op=dist.ReduceOp.MIN, group=group)
        if not int(flag.item()):
            if sg is not None:
                sg.close()
            sg = _SHARED[key] = _SharedGrid(spec, rank, group)
        return sg if sg.ok else None

    # ------------------------------------------------------------------ option 1 / 2
    def _read_locations(self):
        """Where to estimate for `option` 1 (cross-validation: the data themselves) and 2
        (jackknife: the records of `jackfl`), krige3d.py:200-208, 310-332.  The reference
        never reads the file (its loop indexes an empty list); like `datafl`, `jackfl` is a
        Geo-EAS file whose columns are taken by NAME: x, y[, z], then the variable and, for
        ikrige 2/3, the external-drift variable.  Returns x, y, z, true, ext (file order)."""
        if self.koption == 1:
            src, prop = self.data.vr, self.data.property_name
        else:
            jk = SpatialData(self.jackfl) if self._jackknife is None else _MemoryData(self._jackknife)
            src, prop = jk.vr, jk.property_name
        col = lambda nm: np.ascontiguousarray(src[nm], dtype=np.float64)
        x, y, z = col('x'), col('y'), col('z')
        true = col(prop[0]) if len(prop) > 0 else np.full(x.shape, np.nan)
        true = np.where((true < self.tmin) | (true >= self.tmax), np.nan, true)  # krige3d.py:331
        ext = None
        if self.ktype in (2, 3):
            if len(prop) < 2:
                raise ValueError("ikrige 2/3 needs the external variable at every location")
            ext = col(prop[1])
        return x, y, z, true, ext

    def _kt3d_locations(self, neighbours=False):
        """Cross-validation / jackknife: the same two kernels, with the listed locations
        grouped by super block in place of the grid nodes (k3d_krige_points).  Every rank of
        a process group computes all locations (these runs are small).  Locations outside
        the grid extent are searched from the nearest edge super block -- a warning says how
        many (`locations_outside_grid`): inside the grid the template of a location's block
        holds every sample within the radius, outside it need not."""
        t0 = time.time()
        x, y, z, true, ext = self._read_locations()
        n = x.shape[0]
        s = self.searcher
        outside = ((x < self.xmn - 0.5 * self.xsiz) | (x > self.xmn + (self.nx - 0.5) * self.xsiz) |
                   (y < self.ymn - 0.5 * self.ysiz) | (y > self.ymn + (self.ny - 0.5) * self.ysiz) |
                   (z < self.zmn - 0.5 * self.zsiz) | (z > self.zmn + (self.nz - 0.5) * self.zsiz))
        self.locations_outside_grid = int(np.count_nonzero(outside))
        if self.locations_outside_grid:
            # such a location is searched from the nearest edge super block (the reference's
            # getindx has no defined answer for it): samples within the radius of the location
            # but outside that block's template are not seen
            warnings.warn("%d of %d locations lie outside the grid extent; their neighbourhoods "
                          "are those of the nearest edge super block and may miss samples"
                          % (self.locations_outside_grid, n))
        blk = s.block_of_locations(x, y, z)
        order = np.argsort(blk, kind='stable')
        nsup = s.nxsup * s.nysup * s.nzsup
        start = np.concatenate([[0], np.cumsum(np.bincount(blk, minlength=nsup))]).astype(np.int32)
        dev = self._device()
        up = lambda a: None if a is None else torch.from_numpy(
            np.ascontiguousarray(a)).pin_memory().to(dev, non_blocking=True)
        pt = dict(x=up(x[order]), y=up(y[order]), z=up(z[order]),
                  ext=None if ext is None else up(ext[order]), start=up(start))
        d = self._upload(self._host_inputs(), None)
        out = self._alloc_outputs(n, neighbours)
        if n > 0:
            self._launch_points(self._params_struct(), d, pt, n, out)
        res = {}
        for k, v in out.items():
            view, arr = _pinned_array(v.numel(), v.dtype)
            view.copy_(v, non_blocking=True)
            res[k] = arr
        torch.cuda.synchronize(dev)
        inv = np.empty(n, dtype=np.int64)
        inv[order] = np.arange(n)
        self.locations = np.stack([x, y, z], axis=1)
        self.true_values = true
        self.estimation = res['est'][inv]
        self.estimation_variance = res['var'][inv]
        self.status = res['status'][inv]
        if neighbours:
            idx = res['nbr_idx'].reshape(-1, self.ndmax).astype(np.int64)[inv]
            self.neighbour_count = res['nbr_cnt'][inv]
            self.candidates_tested = res['ncand'][inv]
            self.neighbours = np.where(idx >= 0, s.sort_index[np.maximum(idx, 0)], -1)
            self.neighbours_sorted_space = idx
        self.timings['krige_s'] = time.time() - t0
        self.timings['total_s'] = time.time() - t0 + self.timings.get('setup_s', 0.0)
        if self.verbose:
            print("Kriging Finished.")

    def _launch_points(self, P, d, pt, n, out):
        ptr = lambda t: C.c_void_p(None if t is None else t.data_ptr())
        I = self._inputs_struct(d)
        Q = _native.K3dPoints()
        Q.npts = int(n)
        Q.x, Q.y, Q.z, Q.ext, Q.start = ptr(pt['x']), ptr(pt['y']), ptr(pt['z']), ptr(pt['ext']), \
            ptr(pt['start'])
        Q.exclude_coincident = 1  # krige3d.py:359: option != 0
        O = self._outputs_struct(out)
        stream = torch.cuda.current_stream(self._device()).cuda_stream
        _native.check(_native.lib().k3d_krige_points(C.byref(P), C.byref(I), C.byref(Q),
                                                    C.byref(O), C.c_void_p(stream)))

    kriging = kt3d

    def write_output(self, path=None, unest=None):
        """Write `outfl`: Estimate and EstimationVariance, x fastest
        (gslib_help/kt3d_help.md:37-41); for `option` 1/2 one record per location:
        X, Y, Z, True, Estimate, EstimationVariance, Error (estimate - true)."""
        if self.estimation is None:
            raise RuntimeError("run kt3d() first")
        if self.koption != 0:
            write_gslib_validation(path or self.outfl, self.locations, self.true_values,
                                   self.estimation, self.estimation_variance, unest=unest)
            return
        write_gslib_grid(path or self.outfl, self.estimation, self.estimation_variance,
                         unest=unest)

    def view2d(self):
        raise NotImplementedError("plotting is outside the scope of the GPU build")

    def view3d(self):
        pass
