"""Per-step timeline of the factorization schedule: python tools/timeline.py n nb [key=value ...]"""
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
ev = {}
for i in range(cnt):
    ev[(tags[i] // 1000, tags[i] % 1000)] = ms[i]
steps = sorted({k[1] for k in ev})
print(f"n={n} nb={nb} {kv} factor {st.factor_ms:.2f} ms")
print("step | chain: trsm+b1  getrf   idle-before | bulk: trsm   b2+b3a   b3b   idle-before | (ms)")
prev_c = prev_b = None
for k in steps:
    g = lambda kind: ev.get((kind, k))
    t1, t2, t3, t4, t5, t6, t7 = (g(i) for i in range(1, 8))
    f = lambda a, b: f"{(b - a):6.3f}" if a is not None and b is not None else "   -  "
    print(f"{k:3d}  |        {f(t1, t2)}  {f(t2, t3)}   {f(prev_c, t1)}     |     {f(t4, t5)}  {f(t5, t6)}  {f(t6, t7)}   {f(prev_b, t4)}")
    prev_c = t3 if t3 is not None else prev_c
    prev_b = t7 if t7 is not None else prev_b
