"""Single GEMM launch sequence for ncu captures: python tools/one_gemm.py variant M N K [reps]"""
import importlib, os, sys
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
m = importlib.import_module("mixed-precision_lu_factorization_b200")
variant, M, N, K = [int(a) for a in sys.argv[1:5]]
reps = int(sys.argv[5]) if len(sys.argv) > 5 else 3
cm = lambda t: t.t().contiguous().t()
A = cm(torch.randn(M, K, device="cuda").half()); B = cm(torch.randn(K, N, device="cuda").half())
C = cm(torch.zeros(M, N, device="cuda"))
for _ in range(reps):
    m.gemm16(variant, A, B, C, alpha=-1.0, beta=1.0)
print("ok")
