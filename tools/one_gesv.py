"""One (or a few) gesv calls for ncu launch lists: python tools/one_gesv.py n [nb] [reps]"""
import importlib, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
m = importlib.import_module("mixed-precision_lu_factorization_b200")
n = int(sys.argv[1]); nb = int(sys.argv[2]) if len(sys.argv) > 2 else 0; reps = int(sys.argv[3]) if len(sys.argv) > 3 else 1
s = m.Solver(0)
A, b = m.generate(n, seed=1)
opts = m.default_options()
if nb: opts.nb = nb
for _ in range(reps):
    x, st = s.gesv(A, b, opts)
d = st.as_dict()
print({k: d[k] for k in ("iters", "backward_error", "factor_ms", "solve_ms", "kernel_launches", "gemm_launches", "trailing_ms")})
