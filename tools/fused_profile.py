"""Per-step time stamps of the fused GETRF launches (development aid; csrc/getrf_fused.cu).
    python tools/fused_profile.py [n=32768] [nb=2048] [launches=0,5] [key=value options]
Prints, for the chosen fused launches of one factorization, every step (leaf / GEMM step with its tile count and K) with
its duration in clock cycles of CTA 0's SM, and a summary per kind."""
import ctypes as C, importlib, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
m = importlib.import_module("mixed-precision_lu_factorization_b200")
pos = [a for a in sys.argv[1:] if "=" not in a]
kv = dict(a.split("=", 1) for a in sys.argv[1:] if "=" in a)
n = int(pos[0]) if pos else 32768
nb = int(pos[1]) if len(pos) > 1 else 2048
which = [int(x) for x in kv.pop("launches", "0,5").split(",")]
lib = m.load_library()
lib.mplu_debug_fused_profile_enable.argtypes = [C.c_void_p, C.c_int]
lib.mplu_debug_fused_profile.argtypes = [C.c_void_p, C.c_int, C.POINTER(C.c_longlong), C.c_int]
s = m.Solver(0)
A, b = m.generate(n, seed=1)
opts = m.default_options(nb=nb, **{k: int(v) for k, v in kv.items()})
x, st = s.gesv(A, b, opts)
lib.mplu_debug_fused_profile_enable(s._ctx, 1)
for _ in range(2):
    x, st = s.gesv(A, b, opts)
print(f"n={n} nb={nb} {kv} factor {st.factor_ms:.2f} ms solve {st.solve_ms:.2f} ms iters {st.iters} launches {st.kernel_launches}")
buf = (C.c_longlong * (4 * 300))()
for li in which:
    k = lib.mplu_debug_fused_profile(s._ctx, li, buf, 300)
    if k <= 0:
        print(f"launch {li}: not profiled")
        continue
    rec = [tuple(buf[4 * i: 4 * i + 4]) for i in range(k)]
    end_clk, wall_ns = rec[-1][3], rec[-1][2]
    tot = end_clk - rec[0][3]
    print(f"launch {li}: {k - 1} steps, {tot} cycles, {wall_ns / 1e3:.1f} us wall = {tot / max(wall_ns, 1) * 1e3:.0f} MHz")
    summ = {}
    for i in range(k - 1):
        kind, tiles, kk, clk = rec[i]
        dur = (rec[i + 1][3] if i + 1 < k - 1 else end_clk) - clk
        key = "leaf" if kind == 1 else f"gemm K={kk} tiles={tiles}"
        a = summ.setdefault(key, [0, 0])
        a[0] += 1; a[1] += dur
        if "verbose" in kv or k <= 40:
            print(f"   step {i:3d} {key:28s} {dur:8d} cyc")
    for key, (cnt, cyc) in sorted(summ.items(), key=lambda t: -t[1][1]):
        print(f"   {key:28s} x{cnt:3d}  {cyc:9d} cyc  avg {cyc // cnt:7d}  ({100.0 * cyc / tot:4.1f} %)")
