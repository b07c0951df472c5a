"""Block-cyclic solver with all logical ranks on one GPU: python tools/dist_local.py n nb P Q [key=value ...]"""
import importlib, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
m = importlib.import_module("mixed-precision_lu_factorization_b200")
pos = [a for a in sys.argv[1:] if "=" not in a]
kv = dict(a.split("=", 1) for a in sys.argv[1:] if "=" in a)
n, nb, P, Q = (int(a) for a in pos[:4])
ds = m.DistSolver(0, P, Q)
As, bs = ds.generate(n, nb, seed=1)
opts = m.default_options(**{k: int(v) for k, v in kv.items()})
for rep in range(2):
    xs, st = ds.gesv(n, nb, As, bs, opts, allow_noconv=True)
d = st.as_dict()
err = max((x - 1).abs().max().item() for x in xs)
print(f"n={n} nb={nb} grid {P}x{Q} local: total {d['total_ms']:.2f} ms factor {d['factor_ms']:.2f} solve {d['solve_ms']:.2f} iters {d['iters']} "
      f"conv {d['converged']} be {d['backward_error']:.2e} first_be {d['first_backward_error']:.2e} status {d['status_bits']} "
      f"launches {d['kernel_launches']} err {err:.1e}", flush=True)
