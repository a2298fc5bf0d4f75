"""ctypes binding of the C ABI in include/b200_clover.h (the same symbols the Chroma adapter links).

There is no fallback: if the CUDA library is missing or fails to load, importing callers get a loud
B200LibraryError -- nothing in this package can compute on the CPU.
"""
import ctypes as C
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
# B200_LIB_TAG selects a tuning variant built with B200_BUILD_TAG (chroma_b200/build.py); default: the product library
LIB_PATH = os.path.join(_HERE, "libb200clover%s.so" % ("_" + os.environ["B200_LIB_TAG"] if os.environ.get("B200_LIB_TAG") else ""))

B200_SINGLE, B200_DOUBLE = 4, 8
B200_RECONS_NONE, B200_RECONS_12 = 18, 12
B200_SOLVER_CG, B200_SOLVER_BICGSTAB = 0, 1
B200_PLUS, B200_MINUS = 1, -1
B200_PRECOND_ASYMMETRIC, B200_PRECOND_SYMMETRIC = 0, 1
B200_MAX_SHIFTS = 32
B200_OK, B200_ERR_ARG, B200_ERR_CUDA, B200_ERR_STATE, B200_ERR_BREAKDOWN, B200_ERR_COMM = 0, 1, 2, 3, 4, 5


class B200LibraryError(RuntimeError):
    pass


class B200Error(RuntimeError):
    def __init__(self, code, msg):
        super().__init__("b200 error %d: %s" % (code, msg))
        self.code = code


class SolveInfo(C.Structure):
    _fields_ = [("n_count", C.c_int), ("converged", C.c_int), ("resid", C.c_double), ("rel_resid", C.c_double),
                ("rsd_sq_iter", C.c_double), ("secs", C.c_double), ("secs_total", C.c_double), ("gflops", C.c_double),
                ("n_updates", C.c_int), ("reserved", C.c_int)]

    def as_dict(self):
        return {k: getattr(self, k) for k, _ in self._fields_}


ALLGATHER_FN = C.CFUNCTYPE(C.c_int, C.c_void_p, C.c_void_p, C.c_void_p, C.c_size_t)
BARRIER_FN = C.CFUNCTYPE(C.c_int, C.c_void_p)


class Comm(C.Structure):
    _fields_ = [("rank", C.c_int), ("size", C.c_int), ("allgather", ALLGATHER_FN), ("barrier", BARRIER_FN),
                ("user", C.c_void_p)]


# every symbol include/b200_clover.h declares: name -> (restype, argtypes)
_vp, _i, _d = C.c_void_p, C.c_int, C.c_double
_i4 = C.POINTER(C.c_int)
SYMBOLS = {
    "b200_last_error": (C.c_char_p, []),
    "b200_version": (C.c_char_p, []),
    "b200_device_count": (_i, []),
    "b200_create": (_i, [C.POINTER(_vp), _i, _i4, _i4, _i4, C.POINTER(Comm), _i]),
    "b200_destroy": (None, [_vp]),
    "b200_load_gauge": (_i, [_vp, C.POINTER(_vp), _i, C.POINTER(_d), _i, _i]),
    "b200_load_clover": (_i, [_vp, _vp, _vp, _i]),
    "b200_make_clover": (_i, [_vp, _d, _d, _d, _i, _i]),
    "b200_get_clover": (_i, [_vp, _vp, _vp, _i]),
    "b200_clover_logdet": (_i, [_vp, C.POINTER(_d)]),
    "b200_clover_logdet_oo": (_i, [_vp, C.POINTER(_d)]),
    "b200_set_preconditioning": (_i, [_vp, _i]),
    "b200_set_twisted_mass": (_i, [_vp, _d]),
    "b200_dslash": (_i, [_vp, _vp, _vp, _i, _i, _i]),
    "b200_clover_apply": (_i, [_vp, _vp, _vp, _i, _i, _i]),
    "b200_clover_matpc": (_i, [_vp, _vp, _vp, _i, _i]),
    "b200_invert": (_i, [_vp, _vp, _vp, _i, _i, _d, _i, C.POINTER(SolveInfo)]),
    "b200_invert_mdagm": (_i, [_vp, _vp, _vp, _i, _i, _d, _i, C.POINTER(SolveInfo)]),
    "b200_invert_reliable": (_i, [_vp, _vp, _vp, _i, _d, _d, _i, _i, C.POINTER(SolveInfo)]),
    "b200_invert_reliable_bicgstab": (_i, [_vp, _vp, _vp, _i, _d, _d, _i, _i, C.POINTER(SolveInfo)]),
    "b200_invert_multishift": (_i, [_vp, C.POINTER(_vp), _vp, _i, _i, C.POINTER(_d), C.POINTER(_d), _i, C.POINTER(SolveInfo)]),
    "b200_qprop": (_i, [_vp, _vp, _vp, _i, _i, _i, _d, _i, C.POINTER(SolveInfo)]),
    "b200_field_alloc": (_i, [_vp, C.POINTER(_vp)]),
    "b200_field_free": (None, [_vp, _vp]),
    "b200_field_upload": (_i, [_vp, _vp, _vp, _i]),
    "b200_field_download": (_i, [_vp, _vp, _vp, _i]),
    "b200_field_zero": (_i, [_vp, _vp]),
    "b200_mfield_alloc": (_i, [_vp, _i, C.POINTER(_vp)]),
    "b200_mfield_upload": (_i, [_vp, _vp, _i, _vp, _i]),
    "b200_mfield_download": (_i, [_vp, _vp, _i, _vp, _i]),
    "b200_field_nrhs": (_i, [_vp]),
    "b200_dev_dslash": (_i, [_vp, _vp, _vp, _i, _i]),
    "b200_dev_clover_apply": (_i, [_vp, _vp, _vp, _i, _i]),
    "b200_dev_clover_matpc": (_i, [_vp, _vp, _vp, _i]),
    "b200_dev_time_matpc": (_i, [_vp, _vp, _vp, _i, _i, C.POINTER(_d)]),
    "b200_dev_norm2": (_i, [_vp, _vp, C.POINTER(_d)]),
    "b200_dev_inner": (_i, [_vp, _vp, _vp, C.POINTER(_d)]),
    "b200_dev_invert": (_i, [_vp, _vp, _vp, _i, _d, _i, C.POINTER(SolveInfo)]),
    "b200_dev_invert_mdagm": (_i, [_vp, _vp, _vp, _i, _d, _i, C.POINTER(SolveInfo)]),
    "b200_dev_invert_reliable": (_i, [_vp, _vp, _vp, _d, _d, _i, _i, C.POINTER(SolveInfo)]),
    "b200_dev_invert_reliable_bicgstab": (_i, [_vp, _vp, _vp, _d, _d, _i, _i, C.POINTER(SolveInfo)]),
    "b200_dev_invert_multishift": (_i, [_vp, _vp, _vp, _i, C.POINTER(_d), C.POINTER(_d), _i, C.POINTER(SolveInfo)]),
    "b200_dev_iterate_begin": (_i, [_vp, _vp, _vp, _i]),
    "b200_dev_iterate": (_i, [_vp, _i, _i]),
    "b200_dev_time_solver_kernels": (_i, [_vp, _i, _i, C.POINTER(_d), _i, C.POINTER(_i)]),
    "b200_stream": (_vp, [_vp]),
    "b200_sync": (_i, [_vp]),
    "b200_launch_count": (C.c_longlong, [_vp]),
    "b200_host_alloc": (_i, [C.POINTER(_vp), C.c_size_t]),
    "b200_host_free": (None, [_vp]),
    "b200_local_volume": (_i, [_vp, _i4]),
}

_LIB = None


def load():
    """dlopen the engine and bind every declared symbol; raises B200LibraryError if anything is missing."""
    global _LIB
    if _LIB is not None:
        return _LIB
    if not os.path.exists(LIB_PATH):
        raise B200LibraryError("%s not found: build it with `python -m chroma_b200.build` (there is no CPU fallback)" % LIB_PATH)
    try:
        lib = C.CDLL(LIB_PATH)
    except OSError as e:
        raise B200LibraryError("cannot load %s: %s" % (LIB_PATH, e))
    for name, (res, args) in SYMBOLS.items():
        try:
            fn = getattr(lib, name)
        except AttributeError:
            raise B200LibraryError("%s does not export %s" % (LIB_PATH, name))
        fn.restype = res
        fn.argtypes = args
    _LIB = lib
    return lib


def check(rc):
    if rc != 0:
        raise B200Error(rc, load().b200_last_error().decode())
