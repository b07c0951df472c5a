"""One (or a few) gesv calls for timing / ncu launch lists:
    python tools/one_gesv.py n [nb] [reps] [key=value ...]      (keys: any mplu_options field, e.g. lookahead=0 side_sms=32)"""
import importlib, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
m = importlib.import_module("mixed-precision_lu_factorization_b200")
pos = [a for a in sys.argv[1:] if "=" not in a]
kv = dict(a.split("=", 1) for a in sys.argv[1:] if "=" in a)
n = int(pos[0]); nb = int(pos[1]) if len(pos) > 1 else 0; reps = int(pos[2]) if len(pos) > 2 else 1
s = m.Solver(0)
A, b = m.generate(n, seed=1)
opts = m.default_options(**{k: (float(v) if k == "tol" else int(v)) for k, v in kv.items()})
if nb: opts.nb = nb
best = None
for _ in range(reps):
    x, st = s.gesv(A, b, opts)
    d = st.as_dict()
    if best is None or d["total_ms"] < best["total_ms"]: best = d
d = best
tf = 2 / 3 * n ** 3 / (d["total_ms"] * 1e-3) / 1e12
print(f"n={n} nb={opts.nb} {kv} best of {reps}: total {d['total_ms']:.2f} ms = {tf:.1f} TFLOP/s | factor {d['factor_ms']:.2f} solve {d['solve_ms']:.2f} "
      f"iters {d['iters']} be {d['backward_error']:.2e} launches {d['kernel_launches']} gemms {d['gemm_launches']} "
      f"trailing {d['trailing_ms']:.2f} ms ({d['trailing_launches']}) err {(x - 1).abs().max().item():.1e}", flush=True)
