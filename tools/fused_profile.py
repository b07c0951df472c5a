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
lib.mplu_debug_fused_raw.argtypes = [C.c_void_p, C.c_int, C.POINTER(C.c_longlong), C.c_int]
s = m.Solver(0)
A, b = m.generate(n, seed=1)
opts = m.default_options(nb=nb, **{k: int(v) for k, v in kv.items()})
x, st = s.gesv(A, b, opts)
lib.mplu_debug_fused_profile_enable(s._ctx, 2)
for _ in range(2):
    x, st = s.gesv(A, b, opts)
print(f"n={n} nb={nb} {kv} factor {st.factor_ms:.2f} ms solve {st.solve_ms:.2f} ms iters {st.iters} launches {st.kernel_launches}")
buf = (C.c_longlong * (4 * 400))()
for li in which:
    k = lib.mplu_debug_fused_profile(s._ctx, li, buf, 400)
    if k <= 0:
        print(f"launch {li}: not profiled")
        continue
    rec = [tuple(buf[4 * i: 4 * i + 4]) for i in range(k)]
    leafclk = [r[3] for r in rec if r[0] == -2]
    rec = [r for r in rec if r[0] != -2]
    k = len(rec)
    end_clk, wall_ns = rec[-1][3], rec[-1][2]
    tot = end_clk - rec[0][3]
    print(f"launch {li}: {k - 1} steps, {tot} cycles, {wall_ns / 1e3:.1f} us wall = {tot / max(wall_ns, 1) * 1e3:.0f} MHz")
    raw = (C.c_longlong * 640)()
    nraw = lib.mplu_debug_fused_raw(s._ctx, li, raw, 640)
    summ = {}
    prod = {}
    for i in range(k - 1):
        kind, tiles, kk, clk = rec[i]
        sub = ((tiles >> 20) & 0xFFFFF, (tiles >> 40) & 0xFFFFF, (kk >> 20) & 0xFFFFF)  # tfull, epilogue done, barrier entered (CTA 0)
        tiles &= 0xFFFFF; kk &= 0xFFFFF
        dur = (rec[i + 1][3] if i + 1 < k - 1 else end_clk) - clk
        key = "leaf" if kind == 1 else f"gemm K={kk} tiles={tiles}"
        if nraw >= 384 + 4 * i + 4 and kind == 0 and i < 61:
            p = prod.setdefault(key, [0, 0, 0, 0, 0])
            p[0] += 1
            for q in range(4):
                p[1 + q] += max(0, min(raw[384 + 4 * i + q] - clk, 1 << 20))
        a = summ.setdefault(key, [0, 0, 0, 0, 0])
        a[0] += 1; a[1] += dur; a[2] += sub[0]; a[3] += sub[1]; a[4] += sub[2]
        if "verbose" in kv or k <= 40:
            print(f"   step {i:3d} {key:28s} {dur:8d} cyc")
    if leafclk and leafclk[0]:
        names = ["load", "P1+copy", "P2+I1", "P3"] + ["P1+copy", "P2+I1", "P3"] * 2 + ["P1+copy", "tail", "merge", "write-back"]
        sub = leafclk[40:46]
        sub3 = leafclk[46:49]
        if len(sub3) == 3 and sub3[0]:
            print(f"   P3 (kb=1), warp 3: woke up {sub3[0] - sub[3]} after the commit, TMEM->S {sub3[1] - sub3[0]}, barrier {sub3[2] - sub3[1]}")
        leafclk = leafclk[:15]
        if len(sub) == 6 and sub[0]:
            print("   P3 (kb=1) on the tensor cores: " + " ".join(f"{nm}={sub[i + 1] - sub[i]}" for i, nm in enumerate(["stage", "fence+sync", "issue+commit", "wait", "TMEM->S+sync"])))
        d = [leafclk[i + 1] - leafclk[i] for i in range(len(leafclk) - 1) if leafclk[i + 1] > 0]
        print("   last leaf phases (cycles): " + " ".join(f"{nm}={v}" for nm, v in zip(names, d)))
    for key, (cnt, p0, p1, p2, p3) in prod.items():
        print(f"   {key:28s} first tile: producer awake @{p0 // cnt}, after its proxy fence @{p1 // cnt}, first operands landed @{p2 // cnt}, last MMA committed @{p3 // cnt}")
    for key, (cnt, cyc, s0, s1, s2) in sorted(summ.items(), key=lambda t: -t[1][1]):
        print(f"   {key:28s} x{cnt:3d}  {cyc:9d} cyc  avg {cyc // cnt:7d}  ({100.0 * cyc / tot:4.1f} %)   CTA0 avg: accumulator ready @{s0 // cnt}, "
              f"epilogue done @{s1 // cnt}, barrier entered @{s2 // cnt}")
