"""Loader for the in-tree native library (chase_b200/lib/libchase_b200.so).

There is no Python/CPU fallback: if the library is missing the import fails
loudly and tells the user how to build it.
"""
import ctypes
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "lib", "libchase_b200.so")
_lib = None


def build(verbose=False):
    """Compile the CUDA kernels + host layer for sm_100a (nvcc cross-compiles without a GPU)."""
    import subprocess

    out = None if verbose else subprocess.DEVNULL
    subprocess.check_call(["make", "-C", _HERE, "-j2"], stdout=out)
    return LIB_PATH


def lib():
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise RuntimeError(
                f"chase_b200: native library {LIB_PATH} not built. Run `make -C chase_b200` "
                "(or `python -c 'import __graft_entry__ as g; g.build()'`). There is no CPU fallback."
            )
        _lib = ctypes.CDLL(LIB_PATH, mode=ctypes.RTLD_GLOBAL)
        _lib.chase_b200_version.restype = ctypes.c_char_p
        _lib.chase_b200_dmma_peak.restype = ctypes.c_double
        _lib.chase_b200_dmma_peak.argtypes = [ctypes.c_int, ctypes.c_void_p]
        _lib.chase_b200_trsm_ws_bytes.restype = ctypes.c_size_t
        _lib.chase_b200_trsm_ws_bytes.argtypes = [ctypes.c_int64, ctypes.c_int]
        _lib.chase_b200_heev_ws_bytes.restype = ctypes.c_size_t
        _lib.chase_b200_heev_ws_bytes.argtypes = [ctypes.c_int64, ctypes.c_int]
        _lib.chase_b200_hemm_tf32_scratch_bytes.restype = ctypes.c_size_t
        _lib.chase_b200_hemm_tf32_scratch_bytes.argtypes = [ctypes.c_int64, ctypes.c_int64, ctypes.c_int]
        _lib.chase_b200_last_sp_filter_cols_.restype = ctypes.c_double
        _lib.chase_b200_launch_count.restype = ctypes.c_ulonglong
        _lib.chase_b200_trace_copy_.restype = ctypes.c_size_t
        _lib.chase_b200_trace_copy_.argtypes = [ctypes.c_char_p, ctypes.c_size_t]
        _lib.chase_b200_qr_log_copy_.restype = ctypes.c_size_t
        _lib.chase_b200_qr_log_copy_.argtypes = [ctypes.c_char_p, ctypes.c_size_t]
    return _lib
