"""GPU parity tests of the individual kernels, called through the C ABI (include/mplu.h)."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def cm(t):
    """column-major copy (same logical values, stride(0) == 1)"""
    return t.t().contiguous().t()


@pytest.mark.parametrize("variant", [0, 1, 2, 3])
@pytest.mark.parametrize("shape", [(128, 256, 64), (384, 512, 256), (1000, 700, 192), (136, 36, 128)])
@pytest.mark.parametrize("dtype", ["float16", "bfloat16"])
def test_gemm16_matches_fp32_reference(mplu, variant, shape, dtype):
    import torch
    torch.manual_seed(1)
    M, N, K = shape
    dt = getattr(torch, dtype)
    A = cm(torch.randn(M, K, device="cuda").to(dt))
    B = cm(torch.randn(K, N, device="cuda").to(dt))
    C0 = cm(torch.randn(M, N, device="cuda"))
    C = cm(C0.clone())
    kw = dict(alpha=-0.5, beta=1.0, want_shadow=True, hscale=0.25)
    if variant in (2, 3):
        out, H = mplu.gemm16(variant, None, B, C, a_transposed=cm(A.t().contiguous()), **kw)
    else:
        out, H = mplu.gemm16(variant, A, B, C, **kw)
    # plain PyTorch fp32 reference of the same op (inputs are exactly representable; fp32 accumulation)
    ref = C0 - 0.5 * (A.float() @ B.float())
    scale = ref.abs().max().item()
    # tolerance: fp32 accumulation over K terms of magnitude ~1: K * 2^-24 * few
    assert (out - ref).abs().max().item() <= 4e-6 * scale * max(1.0, K / 256) ** 0.5 + 1e-5
    u16 = 2.0 ** -11 if dtype == "float16" else 2.0 ** -8
    assert (H.float() - 0.25 * ref).abs().max().item() <= 1.01 * u16 * 0.25 * scale + 1e-5


def test_gemm16_sm_limit_and_beta0(mplu):
    import torch
    torch.manual_seed(2)
    A = cm(torch.randn(512, 128, device="cuda").half())
    B = cm(torch.randn(128, 768, device="cuda").half())
    ref = A.float() @ B.float()
    for sms in (2, 16, 0):
        out = mplu.gemm16(1, A, B, None, alpha=1.0, beta=0.0, max_sms=sms)
        assert (out - ref).abs().max().item() <= 1e-4


def test_diag_lu128_matches_reference_elimination(mplu, oracle):
    import torch
    lib = mplu.load_library()
    for seed, make in ((5, lambda: oracle.counter_matrix(128, seed=5)),
                       (6, lambda: np.random.default_rng(6).standard_normal((128, 128)) + 12 * np.eye(128))):
        A = make()
        W = cm(torch.tensor(A, dtype=torch.float32, device="cuda"))
        Li = cm(torch.zeros(128, 128, device="cuda"))
        Ui = cm(torch.zeros(128, 128, device="cuda"))
        torch.cuda.synchronize()
        assert lib.mplu_diag_lu128(W.data_ptr(), W.stride(1), Li.data_ptr(), Ui.data_ptr(), None) == 0
        torch.cuda.synchronize()
        ref = oracle.dgetf2_npv(A.astype(np.float32).astype(np.float64))  # dgetf2_native_npv.cu:18-35 in fp64
        got = W.cpu().double().numpy()
        assert np.abs(got - ref).max() <= 2e-6 * np.abs(ref).max()  # fp32 elimination of a 128-block
        L = np.tril(ref, -1) + np.eye(128)
        U = np.triu(ref)
        assert np.abs(Li.cpu().double().numpy() @ L - np.eye(128)).max() <= 1e-5
        assert np.abs(U @ Ui.cpu().double().numpy() - np.eye(128)).max() <= 1e-5


@pytest.mark.parametrize("n", [1, 2, 127, 1000, 4097])
def test_residual_matches_numpy(mplu, n):
    import torch
    lib = mplu.load_library()
    rng = np.random.default_rng(n)
    A = rng.standard_normal((n, n)); x = rng.standard_normal(n); b = rng.standard_normal(n)
    dA = cm(torch.tensor(A, device="cuda")); dx = torch.tensor(x, device="cuda"); db = torch.tensor(b, device="cuda")
    dr = torch.empty(n, dtype=torch.float64, device="cuda"); dn = torch.zeros(2, dtype=torch.float64, device="cuda")
    torch.cuda.synchronize()
    assert lib.mplu_residual(n, dA.data_ptr(), dA.stride(1), dx.data_ptr(), db.data_ptr(), dr.data_ptr(), dn.data_ptr(), None) == 0
    torch.cuda.synchronize()
    ref = b - A @ x
    tol = 64 * np.finfo(np.float64).eps * (np.abs(A) @ np.abs(x) + np.abs(b)).max()
    assert np.abs(dr.cpu().numpy() - ref).max() <= tol
    assert abs(dn[0].item() - np.abs(dr.cpu().numpy()).max()) == 0.0
    assert dn[1].item() == np.abs(x).max()


@pytest.mark.parametrize("n", [1, 130, 1024])
def test_generator_is_bit_identical_to_oracle(mplu, oracle, n):
    A, b = mplu.generate(n, seed=3)
    ref = oracle.counter_matrix(n, seed=3)
    assert np.array_equal(A.cpu().numpy(), ref)
    np.testing.assert_allclose(b.cpu().numpy(), ref.sum(axis=1), rtol=1e-13)


@pytest.mark.parametrize("seed,dominant", [(1, True), (5, True), (3, False)])
def test_fused_leaf_tensor_core_products_are_fp32_accurate(mplu, oracle, seed, dominant):
    """Inside the fused GETRF launch the 128x128 leaf forms its rank-32 Schur updates and the merges of its inverses on
    the tensor cores from three-part bf16 splits (csrc/leaf_tc.cuh).  n = 256 = one fused launch of two leaves: the first
    leaf's LU block and its explicit inverses must be accurate to a few 1e-6 (a one-part bf16 product would be 4e-3, a two-part one 1e-5), i.e.
    the no-pivot elimination of dgetf2_native_npv.cu:18-35 at fp32 level."""
    import ctypes
    import torch
    n = 256
    A = oracle.counter_matrix(n, seed=seed)
    if not dominant:  # general block whose no-pivot LU has no element growth: random + a dominant diagonal
        rng = np.random.default_rng(seed)
        A = rng.standard_normal((n, n)) + 40.0 * np.eye(n)
    b = A.sum(axis=1)
    dA = torch.tensor(A, dtype=torch.float64, device="cuda").t().contiguous().t()
    s = mplu.Solver(0)
    try:
        x, st = s.gesv(dA, torch.tensor(b, device="cuda"), mplu.default_options(nb=256, fuse_w=256, flow_w=0), allow_noconv=True)
        assert st.gemm_launches == 1  # fused: one launch for the whole GETRF
        LU = s.factors(n).cpu().numpy()
        ref = oracle.dgetf2_npv(A[:128, :128])
        scale = np.abs(ref).max()
        assert np.abs(np.triu(LU[:128, :128] - ref)).max() <= 5e-6 * scale
        Lref = np.tril(ref, -1)
        assert np.abs(np.tril(LU[:128, :128], -1) - Lref).max() <= 5e-6 * np.abs(Lref).max()
        lib = mplu.load_library()
        lib.mplu_debug_block_inverses.argtypes = [ctypes.c_void_p, ctypes.c_int, ctypes.c_void_p, ctypes.c_void_p]
        Li = np.zeros((128, 128), dtype=np.float32, order="F")
        Ui = np.zeros((128, 128), dtype=np.float32, order="F")
        assert lib.mplu_debug_block_inverses(s._ctx, 0, Li.ctypes.data, Ui.ctypes.data) == 0
        L = np.tril(ref, -1) + np.eye(128)
        U = np.triu(ref)
        assert np.abs(Li.astype(np.float64) @ L - np.eye(128)).max() <= 5e-5
        assert np.abs(U @ Ui.astype(np.float64) - np.eye(128)).max() <= 5e-5
        assert np.abs(np.triu(Li, 1)).max() == 0 and np.abs(np.tril(Ui, -1)).max() == 0
    finally:
        s.close()
