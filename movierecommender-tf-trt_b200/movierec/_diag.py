"""ctypes binding of libmovierec_b200_diag.so (include/movierec_b200_diag.h): tcgen05 self-tests and probes used by
tests/test_gpu_tc.py and tools/tc_probe.py, tools/tc_rate.py.  Not part of the product path: nothing in movierec
imports this module."""

import ctypes as C
import os

LIB_NAME = "libmovierec_b200_diag.so"
LIB_PATH = os.environ.get("MR_DIAG_LIB_PATH") or os.path.join(os.path.dirname(os.path.abspath(__file__)), LIB_NAME)

_vp, _i32 = C.c_void_p, C.c_int32
SIGNATURES = {
    "mr_diag_last_error": (C.c_char_p, []),
    "mr_tc_gemm_selftest": (C.c_int, [_vp, _vp, _vp, _i32, _i32, _i32, _i32, _i32, _vp]),
    "mr_bf16x3_gemm_selftest": (C.c_int, [_vp, _vp, _vp, _i32, _i32, _i32, _i32, _vp]),
    "mr_tc_probe": (C.c_int, [_vp, _i32, _i32, _i32, _i32, _i32, _vp, _vp]),
    "mr_tc_rate": (C.c_int, [_i32, _i32, _i32, _i32, _i32, _i32, _vp, _i32, _vp]),
}


def _load():
    if not os.path.exists(LIB_PATH):
        raise ImportError("{} not found at {}: build it with `make -C movierecommender-tf-trt_b200/csrc diag`".format(
            LIB_NAME, LIB_PATH))
    lib = C.CDLL(LIB_PATH)
    for name, (res, args) in SIGNATURES.items():
        fn = getattr(lib, name)
        fn.restype = res
        fn.argtypes = args
    return lib


lib = _load()


def check(rc, what):
    if rc != 0:
        msg = lib.mr_diag_last_error()
        raise RuntimeError("{} failed (status {}): {}".format(what, rc, msg.decode("utf-8", "replace") if msg else ""))
