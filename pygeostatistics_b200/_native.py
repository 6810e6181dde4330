"""ctypes binding of libk3d.so (the C ABI declared in include/k3d.h).

There is no CPU fallback: importing this module raises if the CUDA library has not been
built (python pygeostatistics_b200/csrc/build.py, or __graft_entry__.build()).
"""
import ctypes as C
import os

HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("K3D_LIB", os.path.join(HERE, "csrc", "libk3d.so"))

MAXNST, MAXDRIFT, MAXNDMAX, MAXDIS = 4, 9, 64, 128
FLAG_OK_ONE_SAMPLE = 1


class K3dError(RuntimeError):
    pass


class K3dParams(C.Structure):
    _fields_ = [
        ("nx", C.c_int32), ("ny", C.c_int32), ("nz", C.c_int32), ("_pad0", C.c_int32),
        ("xmn", C.c_double), ("ymn", C.c_double), ("zmn", C.c_double),
        ("xsiz", C.c_double), ("ysiz", C.c_double), ("zsiz", C.c_double),
        ("ndmin", C.c_int32), ("ndmax", C.c_int32), ("noct", C.c_int32), ("flags", C.c_int32),
        ("radsqd", C.c_double),
        ("ktype", C.c_int32), ("itrend", C.c_int32), ("idrift", C.c_int32 * MAXDRIFT),
        ("mdt", C.c_int32),
        ("skmean", C.c_double), ("tmin", C.c_double), ("tmax", C.c_double),
        ("nst", C.c_int32), ("it", C.c_int32 * MAXNST), ("_pad2", C.c_int32),
        ("c0", C.c_double), ("cc", C.c_double * MAXNST), ("aa", C.c_double * MAXNST),
        ("rotmat", C.c_double * ((MAXNST + 1) * 9)),
        ("maxcov", C.c_double), ("resc", C.c_double), ("unbias", C.c_double), ("cbb", C.c_double),
        ("bv", C.c_double * MAXDRIFT),
        ("nxsup", C.c_int32), ("nysup", C.c_int32), ("nzsup", C.c_int32), ("ndb", C.c_int32),
        ("brick", C.c_int32 * 3), ("_pad3", C.c_int32),
    ]


class K3dInputs(C.Structure):
    _fields_ = [
        ("nd", C.c_int64),
        ("sx", C.c_void_p), ("sy", C.c_void_p), ("sz", C.c_void_p), ("sv", C.c_void_p),
        ("ssec", C.c_void_p), ("sbstart", C.c_void_p),
        ("nruns", C.c_int32), ("ntmpl", C.c_int32), ("runs", C.c_void_p),
        ("nodex0", C.c_void_p), ("nodey0", C.c_void_p), ("nodez0", C.c_void_p),
        ("dbxyz", C.c_void_p), ("extgrid", C.c_void_p),
    ]


class K3dOutputs(C.Structure):
    _fields_ = [("est", C.c_void_p), ("var", C.c_void_p), ("nbr_idx", C.c_void_p),
                ("nbr_cnt", C.c_void_p), ("status", C.c_void_p), ("ncand", C.c_void_p)]


class K3dPoints(C.Structure):
    _fields_ = [("npts", C.c_int64), ("x", C.c_void_p), ("y", C.c_void_p), ("z", C.c_void_p),
                ("ext", C.c_void_p), ("start", C.c_void_p),
                ("exclude_coincident", C.c_int32), ("_pad", C.c_int32)]


class K3dLaunchInfo(C.Structure):
    _fields_ = [("grid", C.c_int32), ("block", C.c_int32), ("smem_bytes", C.c_int32),
                ("warps_per_cta", C.c_int32), ("stage_capacity", C.c_int32),
                ("list_capacity", C.c_int32), ("launches", C.c_int32),
                ("search_block", C.c_int32), ("search_smem_bytes", C.c_int32), ("_pad", C.c_int32),
                ("search_ms", C.c_float), ("solve_ms", C.c_float)]


# every symbol include/k3d.h declares
SYMBOLS = ("k3d_version", "k3d_last_error", "k3d_krige_slab", "k3d_krige_points",
           "k3d_build_template", "k3d_krige_slab_host", "k3d_krige_points_host", "k3d_last_launch",
           "k3d_measure_fp64_peak", "k3d_release", "k3d_selftest_math")

_lib = None


def lib():
    """The loaded library.  Raises K3dError when it is missing -- never falls back."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise K3dError(
                "CUDA extension %s not built; run `python pygeostatistics_b200/csrc/build.py`"
                % LIB_PATH)
        L = C.CDLL(LIB_PATH)
        L.k3d_version.restype = C.c_char_p
        L.k3d_last_error.restype = C.c_char_p
        L.k3d_krige_slab.restype = C.c_int
        L.k3d_krige_slab.argtypes = [C.POINTER(K3dParams), C.POINTER(K3dInputs), C.c_int32,
                                     C.c_int32, C.POINTER(K3dOutputs), C.c_void_p]
        L.k3d_krige_points.restype = C.c_int
        L.k3d_krige_points.argtypes = [C.POINTER(K3dParams), C.POINTER(K3dInputs),
                                       C.POINTER(K3dPoints), C.POINTER(K3dOutputs), C.c_void_p]
        L.k3d_krige_points_host.restype = C.c_int
        L.k3d_krige_points_host.argtypes = [C.POINTER(K3dParams), C.POINTER(K3dInputs),
                                            C.POINTER(K3dPoints), C.POINTER(K3dOutputs)]
        L.k3d_krige_slab_host.restype = C.c_int
        L.k3d_krige_slab_host.argtypes = [C.POINTER(K3dParams), C.POINTER(K3dInputs), C.c_int32,
                                          C.c_int32, C.POINTER(K3dOutputs)]
        L.k3d_build_template.restype = C.c_int
        L.k3d_build_template.argtypes = [C.c_int32, C.c_int32, C.c_int32, C.c_double,
                                         C.c_double, C.c_double, C.POINTER(C.c_double),
                                         C.c_double, C.c_void_p, C.c_void_p]
        L.k3d_last_launch.restype = C.c_int
        L.k3d_last_launch.argtypes = [C.POINTER(K3dLaunchInfo)]
        L.k3d_measure_fp64_peak.restype = C.c_int
        L.k3d_measure_fp64_peak.argtypes = [C.POINTER(C.c_double), C.POINTER(C.c_double),
                                            C.c_void_p]
        L.k3d_release.restype = C.c_int
        L.k3d_release.argtypes = []
        L.k3d_selftest_math.restype = C.c_int
        L.k3d_selftest_math.argtypes = [C.c_int64, C.POINTER(C.c_uint64), C.POINTER(C.c_uint64),
                                        C.c_void_p]
        _lib = L
    return _lib


def check(rc):
    if rc != 0:
        raise K3dError("k3d error %d: %s" % (rc, lib().k3d_last_error().decode()))


def selftest_math(n, stream=None):
    """(max ulp distance of exp_neg, sqrt_pos, rcp_fast from the CUDA library, violations)."""
    ulp = (C.c_uint64 * 3)()
    bad = C.c_uint64()
    check(lib().k3d_selftest_math(int(n), ulp, C.byref(bad), C.c_void_p(stream)))
    return [int(v) for v in ulp], int(bad.value)


def last_launch():
    info = K3dLaunchInfo()
    check(lib().k3d_last_launch(C.byref(info)))
    return {f[0]: getattr(info, f[0]) for f in K3dLaunchInfo._fields_ if f[0] != "_pad"}
