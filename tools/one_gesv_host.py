"""mplu_gesv_host from pinned host buffers, streamed (left-looking behind the PCIe copy) vs copy-then-solve:
python tools/one_gesv_host.py n [nb] [reps] [key=value ...]"""
import importlib, os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
m = importlib.import_module("mixed-precision_lu_factorization_b200")
pos = [a for a in sys.argv[1:] if "=" not in a]
kv = {k: int(v) for k, v in (a.split("=", 1) for a in sys.argv[1:] if "=" in a)}
n = int(pos[0]); nb = int(pos[1]) if len(pos) > 1 else 0; reps = int(pos[2]) if len(pos) > 2 else 3
s = m.Solver(0)
A, b = m.generate(n, seed=1)
hA = torch.empty(n, n, dtype=torch.float64, pin_memory=True)
hA.copy_(A.t())
hb = b.cpu().pin_memory()
hx = torch.empty(n, dtype=torch.float64, pin_memory=True)
del A
torch.cuda.empty_cache()
for stream_host in (1, 0):
    opts = m.default_options(nb=nb, stream_host=stream_host, **kv)
    s.gesv_host_ptr(n, hA.data_ptr(), n, hb.data_ptr(), hx.data_ptr(), opts)
    best = None
    for _ in range(reps):
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        st = s.gesv_host_ptr(n, hA.data_ptr(), n, hb.data_ptr(), hx.data_ptr(), opts)
        dt = 1e3 * (time.perf_counter() - t0)
        if best is None or dt < best[0]:
            best = (dt, st.as_dict())
    dt, d = best
    print(f"n={n} nb={opts.nb} stream_host={stream_host} {kv}: wall {dt:.2f} ms = {2 / 3 * n ** 3 / dt / 1e9:.1f} TFLOP/s | "
          f"h2d {d['h2d_ms']:.2f} factor {d['factor_ms']:.2f} solve {d['solve_ms']:.2f} d2h {d['d2h_ms']:.3f} total {d['total_ms']:.2f} "
          f"iters {d['iters']} be {d['backward_error']:.2e} err {float((hx - 1).abs().max()):.1e}", flush=True)
s.close()
