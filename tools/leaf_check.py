"""Accuracy of the fused kernel's tensor-core leaf (development aid): n = 256 = two leaves in one fused launch.
Prints the errors of the first leaf's L\\U block and explicit inverses against an fp64 no-pivot LU, for the fused
(tensor-core leaf) and the launch-per-product (fp32 FMA leaf) paths."""
import ctypes, importlib, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np, torch
m = importlib.import_module("mixed-precision_lu_factorization_b200")
lib = m.load_library()
lib.mplu_debug_block_inverses.argtypes = [ctypes.c_void_p, ctypes.c_int, ctypes.c_void_p, ctypes.c_void_p]


def lu_nopiv(A):
    P = A.copy()
    n = P.shape[0]
    for j in range(n - 1):
        P[j + 1:, j] /= P[j, j]
        P[j + 1:, j + 1:] -= np.outer(P[j + 1:, j], P[j, j + 1:])
    return P


n = 256
rng = np.random.default_rng(3)
cases = {"random+12I": rng.standard_normal((n, n)) + 12.0 * np.eye(n)}
A, _ = m.generate(n, seed=1)
cases["dominant"] = A.cpu().numpy()
for name, A in cases.items():
    ref = lu_nopiv(A[:128, :128].astype(np.float64))
    L = np.tril(ref, -1) + np.eye(128); U = np.triu(ref)
    dA = torch.tensor(A, dtype=torch.float64, device="cuda").t().contiguous().t()
    db = torch.tensor(A.sum(axis=1), device="cuda")
    keep = {}
    for fuse in (0, 256):
        s = m.Solver(0)
        x, st = s.gesv(dA, db, m.default_options(nb=256, fuse_w=fuse, flow_w=0), allow_noconv=True)
        LUall = s.factors(n).cpu().numpy()
        refall = lu_nopiv(A.astype(np.float64))
        eall = np.abs(LUall - refall).max() / np.abs(refall).max()
        LU = LUall[:128, :128]
        keep[fuse] = LUall
        Li = np.zeros((128, 128), dtype=np.float32, order="F"); Ui = np.zeros((128, 128), dtype=np.float32, order="F")
        lib.mplu_debug_block_inverses(s._ctx, 0, Li.ctypes.data, Ui.ctypes.data)
        eU = np.abs(np.triu(LU - ref)).max() / np.abs(U).max()
        eL = np.abs(np.tril(LU - ref, -1)).max() / np.abs(np.tril(ref, -1)).max()
        rL = np.abs(Li.astype(np.float64) @ L - np.eye(128)); rU = np.abs(U @ Ui.astype(np.float64) - np.eye(128))
        blk = lambda R: [[f"{R[32*i:32*i+32, 32*j:32*j+32].max():.1e}" for j in range(4)] for i in range(4)]
        print(f"{name:12s} fuse_w={fuse:3d}: |dU|/max|U| {eU:.2e}  |dL|/max|L| {eL:.2e}  |Li L - I| {rL.max():.2e}  |U Ui - I| {rU.max():.2e} "
              f"iters {st.iters} first_be {st.first_backward_error:.2e} whole |dLU|/max {eall:.2e}")
        if fuse:
            print("   |Li L - I| by 32-blocks:", blk(rL)); print("   |U Ui - I| by 32-blocks:", blk(rU))
        s.close()
    D = np.abs(keep[256] - keep[0])
    print("   fused vs launch-per-product, max |diff| by 128-blocks:", [[f"{D[128*i:128*i+128, 128*j:128*j+128].max():.2e}" for j in range(2)] for i in range(2)],
          " vs fp64 LU (fused):", [[f"{np.abs(keep[256] - refall)[128*i:128*i+128, 128*j:128*j+128].max():.2e}" for j in range(2)] for i in range(2)],
          " (launch-per-product):", [[f"{np.abs(keep[0] - refall)[128*i:128*i+128, 128*j:128*j+128].max():.2e}" for j in range(2)] for i in range(2)])
