"""Where a diagonal tile's GETRF spends its time: diag_lu leaves vs the GEMM launches between them, per step:
python tools/leaf_timeline.py n nb [key=value ...]"""
import ctypes, importlib, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
m = importlib.import_module("mixed-precision_lu_factorization_b200")
lib = m.load_library()
pos = [a for a in sys.argv[1:] if "=" not in a]
kv = dict(a.split("=", 1) for a in sys.argv[1:] if "=" in a)
n = int(pos[0]); nb = int(pos[1])
s = m.Solver(0)
lib.mplu_debug_marks_enable.argtypes = [ctypes.c_void_p, ctypes.c_int]
lib.mplu_debug_marks_enable.restype = None
lib.mplu_debug_timeline.argtypes = [ctypes.c_void_p, ctypes.POINTER(ctypes.c_int), ctypes.POINTER(ctypes.c_float), ctypes.c_int]
lib.mplu_debug_marks_enable(s._ctx, 1)
A, b = m.generate(n, seed=1)
opts = m.default_options(**{k: int(v) for k, v in kv.items()})
opts.nb = nb
for _ in range(3):
    x, st = s.gesv(A, b, opts)
tags = (ctypes.c_int * 4096)(); ms = (ctypes.c_float * 4096)()
cnt = lib.mplu_debug_timeline(s._ctx, tags, ms, 4096)
ev = {tags[i]: ms[i] for i in range(cnt)}
per = nb // 128
print(f"n={n} nb={nb} {kv} factor {st.factor_ms:.2f} ms; per tile: GETRF ms | leaves: sum ms, min/avg/max us | GEMM launches between them: ms")
for t in range((n + nb - 1) // nb):
    g0, g1 = ev.get(2000 + t), ev.get(3000 + t)
    leaf = [1e3 * (ev[9000 + b_] - ev[8000 + b_]) for b_ in range(t * per, (t + 1) * per) if 9000 + b_ in ev and 8000 + b_ in ev]
    if not leaf:
        continue
    tot = (g1 - g0) if g0 is not None and g1 is not None else float("nan")
    if t == 0 and tot != tot:
        tot = ev[9000 + per - 1] - ev[8000]
    print(f"{t:3d} | {tot:6.3f} | {sum(leaf) / 1e3:6.3f}  {min(leaf):5.1f}/{sum(leaf) / len(leaf):5.1f}/{max(leaf):5.1f} | {tot - sum(leaf) / 1e3:6.3f}")
