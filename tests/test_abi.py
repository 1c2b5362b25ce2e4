"""CPU-side checks of the drop-in boundary: the native library loads and exports every
symbol that include/*.h declares (no compute calls here: there is no GPU)."""
import ctypes
import os
import re
import subprocess

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
INC = os.path.join(ROOT, "include")


def _declared(header):
    src = subprocess.check_output(["gcc", "-E", "-P", os.path.join(INC, header)], text=True)
    names = set(re.findall(r"\b((?:chase_b200_\w+)|(?:p?[sdcz]chase_\w*_)|(?:chase_\w+_))\s*\(", src))
    return sorted(names)


@pytest.fixture(scope="module")
def native():
    from chase_b200 import LIB_PATH, build

    if not os.path.exists(LIB_PATH):
        build()
    return ctypes.CDLL(LIB_PATH)


@pytest.mark.parametrize("header", ["chase_b200_kernels.h", "chase_c_interface.h", "chase_b200_comm.h"])
def test_every_declared_symbol_is_exported(native, header):
    names = _declared(header)
    assert len(names) > (20 if header != "chase_b200_comm.h" else 8)
    missing = [n for n in names if not hasattr(native, n)]
    assert not missing, f"declared in {header} but not exported: {missing}"


def test_kernel_api_has_all_four_types(native):
    names = _declared("chase_b200_kernels.h")
    for op in ["gemm", "hemm", "potrf", "trsm", "heev", "colnorms", "gemv_conjt", "lanczos_step"]:
        for s in "sdcz":
            assert f"chase_b200_{op}_{s}" in names


def test_version_and_feature_queries(native):
    native.chase_b200_version.restype = ctypes.c_char_p
    assert b"sm_100a" in native.chase_b200_version()
    flag = ctypes.c_int(-1)
    native.chase_has_cuda_(ctypes.byref(flag))
    assert flag.value == 1
    native.chase_has_mpi_(ctypes.byref(flag))
    assert flag.value == 0


def test_workspace_queries(native):
    native.chase_b200_heev_ws_bytes.restype = ctypes.c_size_t
    native.chase_b200_heev_ws_bytes.argtypes = [ctypes.c_int64, ctypes.c_int]
    assert native.chase_b200_heev_ws_bytes(100, 0) >= 2 * 100 * 100 * 8
    assert native.chase_b200_heev_ws_bytes(100, 1) >= 2 * 100 * 100 * 16
    native.chase_b200_trsm_ws_bytes.restype = ctypes.c_size_t
    native.chase_b200_trsm_ws_bytes.argtypes = [ctypes.c_int64, ctypes.c_int]
    assert native.chase_b200_trsm_ws_bytes(200, 8) == 2 * 128 * 128 * 8


def test_solver_refuses_to_run_without_gpu():
    """No CPU fallback: constructing the backend without a CUDA device must fail loudly."""
    import numpy as np
    import torch

    if torch.cuda.is_available():
        pytest.skip("GPU present")
    code = (
        "import numpy as np, chase_b200\n"
        "H=np.eye(64,order='F')\n"
        "try:\n"
        "    chase_b200.ChASE(H,4,4)\n"
        "except Exception as e:\n"
        "    print('RAISED', type(e).__name__)\n"
    )
    out = subprocess.run(["python", "-c", code], cwd=ROOT, capture_output=True, text=True)
    # a C++ exception crossing the C ABI terminates the process; either way it must not succeed silently
    assert "RAISED" in out.stdout or out.returncode != 0


def test_every_function_of_the_reference_c_interface_is_exported(native):
    """Drop-in check: all 94 functions the reference declares in interface/chase_c_interface.h (names recorded in
    tests/golden/reference_c_symbols.txt) exist in libchase_b200.so."""
    path = os.path.join(ROOT, "tests", "golden", "reference_c_symbols.txt")
    names = [ln.strip() for ln in open(path) if ln.strip() and not ln.startswith("#")]
    assert len(names) >= 90
    missing = [n for n in names if not hasattr(native, n)]
    assert not missing, missing
