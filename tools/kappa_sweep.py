"""Condition-number sweep (BASELINE.json configs[4]): SPD A = H diag(sigma) H, H = I - 2uu^T, sigma geometric in
[1/kappa, 1] (the oracle's spd_kappa_matrix, built here on the device in O(n^2)); b = A*1.
    python tools/kappa_sweep.py [n=16384] [key=value ...]
Prints, per kappa and operand type: refinement iterations, converged flag, fp64 backward error; and the backward error
of an fp64 partial-pivoting LU solve (cuSOLVER through torch) standing in for "the reference's fp64 factors + dgetrs"."""
import importlib, json, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
m = importlib.import_module("mixed-precision_lu_factorization_b200")
pos = [a for a in sys.argv[1:] if "=" not in a]
kv = dict(a.split("=", 1) for a in sys.argv[1:] if "=" in a)
n = int(pos[0]) if pos else 16384
s = m.Solver(0)


spd = m.generate_spd  # symmetric: row-major storage == column-major storage


rows = []
for e in range(2, 9):
    kappa = 10.0 ** e
    A = spd(n, kappa)
    b = A.sum(dim=1)
    rec = dict(kappa=kappa)
    for name, prec in (("fp16", 0), ("bf16", 1)):
        for refine in ([0, 1] if "gmres" in kv else [0]):
            o = {k: int(v) for k, v in kv.items() if k != "gmres"}
            opts = m.default_options(precision=prec, refinement=refine, **{"fp64_fallback": 0, **o})
            x, st = s.gesv(A.t(), b, opts, allow_noconv=True)  # symmetric: the transposed view is the column-major one
            tag = name + ("_gmres" if refine else "")
            rec[tag] = dict(iters=st.iters, converged=st.converged, backward_error=st.backward_error, status=st.status_bits, gmres_iters=st.gmres_iters,
                            max_err=float((x - 1).abs().max().item()), ms=st.total_ms)
    xr = torch.linalg.solve(A, b)
    r = b - A @ xr
    rec["fp64_lu"] = dict(backward_error=float(r.abs().max() / (A.abs().sum(dim=1).max() * xr.abs().max() + b.abs().max())))
    rows.append(rec)
    print(json.dumps(rec), flush=True)
