"""Time stamps of one dataflow GETRF launch (development aid; csrc/getrf_flow.cu).
    python tools/flow_profile.py [n=32768] [nb=2048] [launch=5] [key=value options]
Prints, for the chosen dataflow launch of one factorization: every leaf (start, duration, gap to the previous leaf's end),
what filled each gap (the three A-group tasks between two leaves: taken / dependencies met / signalled, relative to the
previous leaf's end), the helpers' busy time per task kind and the tail after the last leaf."""
import ctypes as C, importlib, os, struct, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
m = importlib.import_module("mixed-precision_lu_factorization_b200")
pos = [a for a in sys.argv[1:] if "=" not in a]
kv = dict(a.split("=", 1) for a in sys.argv[1:] if "=" in a)
n = int(pos[0]) if pos else 32768
nb = int(pos[1]) if len(pos) > 1 else 2048
launch = int(kv.pop("launch", "5"))
lib = m.load_library()
lib.mplu_debug_flow_profile_enable.argtypes = [C.c_void_p, C.c_int]
lib.mplu_debug_flow_profile.argtypes = [C.c_void_p, C.POINTER(C.c_longlong), C.c_int, C.c_char_p, C.c_int, C.POINTER(C.c_int)]
s = m.Solver(0)
A, b = m.generate(n, seed=1)
kw = {k: int(v) for k, v in kv.items()}
kw.setdefault("flow_w", nb)
opts = m.default_options(nb=nb, **kw)
x, st = s.gesv(A, b, opts)
lib.mplu_debug_flow_profile_enable(s._ctx, launch)
for _ in range(2):
    x, st = s.gesv(A, b, opts)
print(f"n={n} nb={nb} {kw} factor {st.factor_ms:.2f} ms solve {st.solve_ms:.2f} ms iters {st.iters} be {st.backward_error:.2e} launches {st.kernel_launches}")
MAXS, MAXT = 1 << 20, 8 << 20
stamps = (C.c_longlong * MAXS)()
tasks = C.create_string_buffer(MAXT)
nl = C.c_int(0)
nt = lib.mplu_debug_flow_profile(s._ctx, stamps, MAXS, tasks, MAXT, C.byref(nl))
if nt < 0:
    print("not profiled", nt)
    sys.exit(0)
L = nl.value
leaf = [(stamps[2 * i], stamps[2 * i + 1]) for i in range(L)]
t0 = leaf[0][0]
KIND = ["P_L", "P_U", "S", "T1", "T2", "X", "Y"]
recs = []
for i in range(nt):
    f = struct.unpack_from("<HBBBB4H4H3H2H", tasks.raw, 32 * i)
    kind, step = f[-2], f[-1]
    g, r, sg, cta = (stamps[2 * L + 4 * i + j] for j in range(4))
    recs.append(dict(i=i, kind=kind, step=step, mt=f[1], nt=f[2], grab=g, ready=r, sig=sg, cta=cta))
print(f"launch {launch}: {L} leaves, {nt} tasks; first leaf start -> last leaf end {(leaf[-1][1] - t0) / 1e3:.1f} us")
for i, (a, e) in enumerate(leaf):
    gap = (a - leaf[i - 1][1]) / 1e3 if i else 0.0
    line = f"  leaf {i:2d}: start {(a - t0) / 1e3:8.1f} us  dur {(e - a) / 1e3:6.1f} us  gap {gap:6.1f} us"
    if i:
        crit = [r for r in recs if r["step"] == i - 1 and ((r["kind"] in (0, 1) and r["mt"] == 0 and r["nt"] == 0) or (r["kind"] == 2 and r["mt"] == 0 and r["nt"] == 0))]
        pe = leaf[i - 1][1]
        line += " | " + "  ".join(f"{KIND[r['kind']]}: taken {(r['grab'] - pe) / 1e3:+.1f} ready {(r['ready'] - pe) / 1e3:+.1f} done {(r['sig'] - pe) / 1e3:+.1f}" for r in crit)
    print(line)
end_all = max(r["sig"] for r in recs)
print(f"  tail after the last leaf: {(end_all - leaf[-1][1]) / 1e3:.1f} us; whole launch {(end_all - t0) / 1e3:.1f} us")
for k in range(7):
    rs = [r for r in recs if r["kind"] == k]
    if not rs:
        continue
    busy = sum(r["sig"] - r["ready"] for r in rs) / 1e3
    waitt = sum(r["ready"] - r["grab"] for r in rs) / 1e3
    print(f"  {KIND[k]:3s}: {len(rs):5d} tasks, dependencies-met -> signalled {busy / len(rs):6.2f} us avg, taken -> dependencies met {waitt / len(rs):7.2f} us avg")
ctas = sorted(set(r["cta"] for r in recs))
print("  tasks per helper CTA:", {c: sum(1 for r in recs if r["cta"] == c) for c in ctas})
