"""tcgen05 building block: the single-tile tensor-core GEMM (operand layouts, descriptors, TMEM)."""

import ctypes as C

import numpy as np
import pytest

pytestmark = pytest.mark.gpu
torch = pytest.importorskip("torch")

CASES = [
    # N, K, a_mn, b_mn
    (128, 32, 0, 0), (64, 64, 0, 0), (16, 8, 0, 0), (256, 32, 0, 0),
    (128, 32, 1, 0), (128, 32, 0, 1), (64, 16, 1, 1), (128, 64, 1, 1), (32, 40, 0, 1), (256, 32, 1, 1), (96, 8, 1, 1),
]


def run_case(N, K, a_mn, b_mn, three_x, seed=0):
    from movierec import _diag as nat
    rng = np.random.default_rng(seed)
    A = rng.normal(size=(128, K)).astype(np.float32)
    B = rng.normal(size=(N, K)).astype(np.float32)
    a_src = np.ascontiguousarray(A.T) if a_mn else A
    b_src = np.ascontiguousarray(B.T) if b_mn else B
    dA, dB = torch.from_numpy(a_src).cuda(), torch.from_numpy(b_src).cuda()
    D = torch.full((128, N), float("nan"), dtype=torch.float32, device="cuda")
    st = C.c_void_p(torch.cuda.current_stream().cuda_stream)
    nat.check(nat.lib.mr_tc_gemm_selftest(C.c_void_p(dA.data_ptr()), C.c_void_p(dB.data_ptr()), C.c_void_p(D.data_ptr()),
                                          N, K, a_mn, b_mn, three_x, st), "mr_tc_gemm_selftest")
    torch.cuda.synchronize()
    want = A.astype(np.float64) @ B.astype(np.float64).T
    got = D.cpu().numpy().astype(np.float64)
    scale = np.abs(A).astype(np.float64) @ np.abs(B).astype(np.float64).T
    return float(np.max(np.abs(got - want) / scale)), got, want


@pytest.mark.parametrize("case", CASES, ids=lambda c: "N{}K{}a{}b{}".format(*c))
def test_tc_gemm_3xtf32_is_fp32_accurate(case):
    if not torch.cuda.is_available():
        pytest.skip("needs a CUDA device")
    err, got, want = run_case(*case, three_x=1)
    assert err < 2e-6, "3xTF32 relative error {:.3e} (first row got {} want {})".format(err, got[0, :4], want[0, :4])


@pytest.mark.parametrize("case", CASES, ids=lambda c: "N{}K{}a{}b{}".format(*c))
def test_tc_gemm_fast_activation_split_is_fp32_accurate(case):
    """three_x = 2: A split by truncation (hi = x & ~0x1fff, lo = x - hi, not re-rounded), B split exactly --
    the combination the production kernels use (activations x pre-packed weights)."""
    if not torch.cuda.is_available():
        pytest.skip("needs a CUDA device")
    err, got, want = run_case(*case, three_x=2)
    assert err < 4e-6, "fast-split 3xTF32 relative error {:.3e}".format(err)


@pytest.mark.parametrize("case", CASES[:2] + CASES[4:6], ids=lambda c: "N{}K{}a{}b{}".format(*c))
def test_tc_gemm_plain_tf32_is_tf32_accurate(case):
    if not torch.cuda.is_available():
        pytest.skip("needs a CUDA device")
    err, _, _ = run_case(*case, three_x=0)
    assert 1e-6 < err < 2e-3, "single-pass TF32 relative error {:.3e}".format(err)


# ---- the operand form of the fused train kernel (tc_fused.cu): 3 x bf16 parts, SWIZZLE_128B tiles ----------------------
BF16_CASES = [
    # N, K, a_mn, b_mn
    (64, 128, 0, 1),   # forward of the fused kernel:  H1 (K-major) x W2 (MN-major)
    (128, 64, 0, 0),   # backward:                     dZ2 (K-major) x W2 (K-major)
    (64, 128, 1, 1),   # weight gradient:              H1^T (MN-major) x dZ2 (MN-major)
    (64, 16, 0, 0), (128, 128, 0, 0), (128, 48, 1, 0), (64, 32, 1, 1), (256, 64, 0, 1), (16, 16, 0, 0),
]


@pytest.mark.parametrize("case", BF16_CASES, ids=lambda c: "N{}K{}a{}b{}".format(*c))
def test_tc_gemm_bf16x3_is_fp32_accurate(case):
    if not torch.cuda.is_available():
        pytest.skip("needs a CUDA device")
    from movierec import _diag as nat
    N, K, a_mn, b_mn = case
    rng = np.random.default_rng(7)
    A = (rng.normal(size=(128, K)) * np.exp(rng.normal(size=(128, K)) * 3)).astype(np.float32)  # wide dynamic range
    B = rng.normal(size=(N, K)).astype(np.float32)
    a_src = np.ascontiguousarray(A.T) if a_mn else A
    b_src = np.ascontiguousarray(B.T) if b_mn else B
    dA, dB = torch.from_numpy(a_src).cuda(), torch.from_numpy(b_src).cuda()
    D = torch.full((128, N), float("nan"), dtype=torch.float32, device="cuda")
    st = C.c_void_p(torch.cuda.current_stream().cuda_stream)
    nat.check(nat.lib.mr_bf16x3_gemm_selftest(C.c_void_p(dA.data_ptr()), C.c_void_p(dB.data_ptr()), C.c_void_p(D.data_ptr()),
                                              N, K, a_mn, b_mn, st), "mr_bf16x3_gemm_selftest")
    torch.cuda.synchronize()
    want = A.astype(np.float64) @ B.astype(np.float64).T
    got = D.cpu().numpy().astype(np.float64)
    scale = np.abs(A).astype(np.float64) @ np.abs(B).astype(np.float64).T
    err = float(np.max(np.abs(got - want) / scale))
    assert err < 2e-6, "bf16x3 relative error {:.3e} (first row got {} want {})".format(err, got[0, :4], want[0, :4])
