"""One or a few calls of the drop-in MPF() (include/MPF.h) from a host buffer, timed as benchmark.cpp:219-222 times it:
    python tools/one_mpf.py n [r=32] [reps=2] [rand]      (rand: a general matrix that really pivots instead of the dominant one)"""
import importlib, os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np
import torch
m = importlib.import_module("mixed-precision_lu_factorization_b200")
pos = [a for a in sys.argv[1:] if a != "rand"]
n = int(pos[0]); r = int(pos[1]) if len(pos) > 1 else 32; reps = int(pos[2]) if len(pos) > 2 else 2
if "rand" in sys.argv:
    hA0 = np.asfortranarray(np.random.default_rng(1).standard_normal((n, n)))
else:
    A, _ = m.generate(n, seed=1, with_rhs=False)
    hA0 = np.asfortranarray(A.t().cpu().numpy().T)
    del A
    torch.cuda.empty_cache()
best = None
for _ in range(reps):
    hA = hA0.copy(order="F")
    t = time.perf_counter()
    ipiv = m.MPF(hA, r)
    dt = time.perf_counter() - t
    best = dt if best is None else min(best, dt)
tf = 2 / 3 * n ** 3 / best / 1e12
swaps = int((ipiv != np.arange(1, n + 1)).sum())
msg = f"MPF n={n} r={r}: best of {reps}: {1e3 * best:.1f} ms = {tf:.2f} TFLOP/s, {swaps} rows swapped"
if n <= 8192:  # P A = L U check against the input
    L = np.tril(hA, -1) + np.eye(n); U = np.triu(hA)
    PA = hA0.copy()
    for j in range(n):
        p = ipiv[j] - 1
        if p != j: PA[[j, p], :] = PA[[p, j], :]
    msg += f", |PA - LU|/|A| = {np.abs(PA - L @ U).max() / np.abs(hA0).max():.2e}"
print(msg, flush=True)
