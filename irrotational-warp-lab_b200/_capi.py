"""ctypes binding of ``libiwb200.so`` (declarations mirror ``include/iwb200.h``).

The library is the product: if it is missing or fails to load there is NO fallback --
importing this module raises, and so does every evaluator of the ``b200`` backend.
"""
from __future__ import annotations

import ctypes as C
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
# The product library.  A/B timing of kernel variants (scripts/ab2.sh) loads another build of it from this directory, named
# by IWB_LIB -- honoured only together with IWB_ALLOW_LIB_OVERRIDE=1, so that a stray environment variable cannot swap the
# library under the tests or the benchmark (bench.py refuses to run with an override).
_override = os.environ.get("IWB_LIB") if os.environ.get("IWB_ALLOW_LIB_OVERRIDE") == "1" else None
LIB_PATH = os.path.join(_HERE, os.path.basename(_override or "libiwb200.so"))
LIB_IS_OVERRIDDEN = _override is not None

MAX_SHELLS = 8
FLUSH_PLANES = 16
TILE_CHAINS = 16
ROW_WIDTH = 6 + 3 * MAX_SHELLS
PEER_HANDLE_BYTES = 64

OK, ERR_INVALID, ERR_GRID_TOO_SMALL, ERR_NO_DEVICE, ERR_CUDA, ERR_NOMEM = range(6)
F64, F32_REF = 0, 1
MODE_EXACT, MODE_FAST = 0, 1

_dp = C.POINTER(C.c_double)


class Params3D(C.Structure):
    _fields_ = [
        ("dtype", C.c_int32), ("mode", C.c_int32),
        ("n0", C.c_int32), ("n1", C.c_int32), ("n2", C.c_int32),
        ("i_begin", C.c_int32), ("i_end", C.c_int32),
        ("x", _dp), ("y", _dp), ("z", _dp),
        ("dx", C.c_double), ("dy", C.c_double), ("dz", C.c_double),
        ("rho", C.c_double), ("sigma", C.c_double), ("v", C.c_double), ("eps", C.c_double),
        ("sinh_rs", C.c_double), ("cosh_rs", C.c_double),
        ("nshell", C.c_int32),
        ("shell_r", C.c_double * MAX_SHELLS),
        ("rho_out", C.c_void_p),
        ("rho_out_is_device", C.c_int32),
        ("tile_stride", C.c_int32), ("tile_first", C.c_int32),
    ]


class Result(C.Structure):
    _fields_ = [
        ("e_pos", C.c_double), ("e_neg", C.c_double),
        ("cnt_pos", C.c_int64), ("cnt_neg", C.c_int64), ("cnt_zero", C.c_int64), ("cnt_nan", C.c_int64),
        ("shell_pos", C.c_double * MAX_SHELLS), ("shell_neg", C.c_double * MAX_SHELLS),
        ("shell_cnt", C.c_int64 * MAX_SHELLS),
        ("kernel_ms", C.c_double), ("total_ms", C.c_double),
    ]


class ParamsAxisym(C.Structure):
    _fields_ = [
        ("dtype", C.c_int32), ("mode", C.c_int32),
        ("nx", C.c_int32), ("ny", C.c_int32),
        ("x", _dp), ("y", _dp),
        ("dx", C.c_double), ("dy", C.c_double),
        ("rho", C.c_double), ("sigma", C.c_double), ("v", C.c_double), ("eps", C.c_double),
        ("sinh_rs", C.c_double), ("cosh_rs", C.c_double),
        ("nshell", C.c_int32),
        ("shell_r", C.c_double * MAX_SHELLS),
        ("rho_out", _dp),
    ]


class ParamsSlice(C.Structure):
    _fields_ = [
        ("n", C.c_int32),
        ("x", _dp), ("y", _dp),
        ("dx", C.c_double), ("dy", C.c_double),
        ("rho", C.c_double), ("sigma", C.c_double), ("v", C.c_double), ("eps", C.c_double),
        ("phi", _dp), ("beta_x", _dp), ("beta_y", _dp), ("rho_adm", _dp),
    ]


class SliceResultC(C.Structure):
    _fields_ = [
        ("pos", C.c_double), ("neg", C.c_double), ("net", C.c_double),
        ("cnt_pos", C.c_int64), ("cnt_neg", C.c_int64), ("cnt_zero", C.c_int64),
        ("kernel_ms", C.c_double), ("total_ms", C.c_double),
    ]


# every symbol include/iwb200.h declares: name -> (restype, argtypes)
class Profile3D(C.Structure):
    _fields_ = [("n_bins", C.c_int32), ("edges", C.POINTER(C.c_double)), ("sum", C.POINTER(C.c_double)),
                ("count", C.POINTER(C.c_int64))]


class ParamsEinstein(C.Structure):
    _fields_ = [("nx", C.c_int32), ("ny", C.c_int32),
                ("beta_x", C.POINTER(C.c_double)), ("beta_y", C.POINTER(C.c_double)),
                ("dx", C.c_double), ("dy", C.c_double), ("imag_tol", C.c_double),
                ("eig_re", C.POINTER(C.c_double)), ("eig_im", C.POINTER(C.c_double)),
                ("eig_max", C.POINTER(C.c_double)), ("ricci_scalar", C.POINTER(C.c_double)),
                ("is_type_i", C.POINTER(C.c_uint8))]


class EinsteinResultC(C.Structure):
    _fields_ = [("cnt_type_i", C.c_int64), ("kernel_ms", C.c_double), ("total_ms", C.c_double)]


SYMBOLS = {
    "iwb_version": (C.c_int, []),
    "iwb_device_count": (C.c_int, []),
    "iwb_last_error": (C.c_char_p, []),
    "iwb_create": (C.c_int, [C.c_int, C.POINTER(C.c_void_p)]),
    "iwb_destroy": (C.c_int, [C.c_void_p]),
    "iwb_device_info": (C.c_int, [C.c_void_p, C.POINTER(C.c_int), C.POINTER(C.c_int), C.POINTER(C.c_int),
                                  C.c_char_p, C.c_int]),
    "iwb_energy3d": (C.c_int, [C.c_void_p, C.POINTER(Params3D), C.POINTER(Result)]),
    "iwb_energy3d_slab_async": (C.c_int, [C.c_void_p, C.POINTER(Params3D), C.c_void_p, C.c_void_p]),
    "iwb_energy3d_tiles_async": (C.c_int, [C.c_void_p, C.POINTER(Params3D), C.c_void_p, C.c_void_p]),
    "iwb_peer_create": (C.c_int, [C.c_void_p, C.c_int, C.c_int, C.c_int, C.POINTER(C.c_void_p), C.c_void_p]),
    "iwb_peer_connect": (C.c_int, [C.c_void_p, C.c_void_p]),
    "iwb_peer_destroy": (C.c_int, [C.c_void_p]),
    "iwb_debug_peer_connect_local": (C.c_int, [C.c_void_p, C.POINTER(C.c_void_p)]),
    "iwb_energy3d_tiles_exchange_async": (C.c_int, [C.c_void_p, C.c_void_p, C.POINTER(Params3D), C.c_void_p, C.c_void_p]),
    "iwb_read_row": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int, C.c_void_p, C.POINTER(Result)]),
    "iwb_reduce_chains": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_void_p, C.POINTER(Result)]),
    "iwb_reduce_table_async": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int, C.c_void_p, C.c_void_p]),
    "iwb_reduce_chains_async": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p]),
    "iwb_plan3d_create": (C.c_int, [C.c_void_p, C.POINTER(Params3D), C.POINTER(C.c_void_p)]),
    "iwb_plan3d_run_async": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p]),
    "iwb_plan3d_run_exchange_async": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]),
    "iwb_plan3d_destroy": (C.c_int, [C.c_void_p]),
    "iwb_reduce_table": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_void_p, C.POINTER(Result)]),
    "iwb_radial_profile3d": (C.c_int, [C.c_void_p, C.POINTER(Params3D), C.POINTER(Profile3D), C.POINTER(Result)]),
    "iwb_energy3d_batch": (C.c_int, [C.c_void_p, C.c_int, C.POINTER(Params3D), C.POINTER(Result)]),
    "iwb_energy_axisym": (C.c_int, [C.c_void_p, C.POINTER(ParamsAxisym), C.POINTER(Result)]),
    "iwb_slice_z0": (C.c_int, [C.c_void_p, C.POINTER(ParamsSlice), C.POINTER(SliceResultC)]),
    "iwb_einstein_z0": (C.c_int, [C.c_void_p, C.POINTER(ParamsEinstein), C.POINTER(EinsteinResultC)]),
    "iwb_measure_fma_peak": (C.c_int, [C.c_void_p, C.c_int, C.POINTER(C.c_double), C.POINTER(C.c_double)]),
    "iwb_selftest_divsqrt": (C.c_int, [C.c_void_p, C.c_uint64, C.c_uint64, C.c_int, C.POINTER(C.c_uint64),
                                       C.POINTER(C.c_uint64), C.POINTER(C.c_uint64)]),
    "iwb_selftest_phi": (C.c_int, [C.c_void_p, C.c_double, C.c_double, C.c_double, C.c_double, C.c_double, C.c_double,
                                   C.c_uint64, C.c_uint64, C.POINTER(C.c_uint64), C.POINTER(C.c_uint64)]),
    "iwb_debug_spans": (C.c_int, [C.c_longlong, C.c_int, C.c_int, C.POINTER(C.c_int), C.c_int]),
    "iwb_debug_tiles_k": (C.c_int, [C.c_int, C.POINTER(C.c_int), C.c_int]),
    "iwb_debug_tiles_j": (C.c_int, [C.c_int, C.POINTER(C.c_int), C.c_int]),
    "iwb_debug_profile_path": (C.c_int, [C.c_void_p]),
    "iwb_selftest_constdiv": (C.c_int, [C.c_void_p, C.c_double, C.c_uint64, C.c_uint64, C.POINTER(C.c_uint64)]),
}

_lib = None


def load():
    """dlopen the in-tree library; raises RuntimeError when it has not been built."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise RuntimeError(
            f"{LIB_PATH} is missing: build it with `python irrotational-warp-lab_b200/build.py` "
            "(the b200 backend has no CPU fallback)")
    lib = C.CDLL(LIB_PATH)
    for name, (res, args) in SYMBOLS.items():
        fn = getattr(lib, name)       # AttributeError if the library does not export it
        fn.restype = res
        fn.argtypes = args
    _lib = lib
    return lib


def last_error():
    return load().iwb_last_error().decode("utf-8", "replace")


def check(rc):
    """Map status codes to the reference's exception types
    (reproduce_rodal_exact.py:25-30; numpy.gradient's ValueError)."""
    if rc == OK:
        return
    msg = last_error()
    if rc in (ERR_INVALID, ERR_GRID_TOO_SMALL):
        raise ValueError(msg)
    if rc == ERR_NOMEM:
        raise MemoryError(msg)
    raise RuntimeError(msg)
