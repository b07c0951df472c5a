"""Per-launch device time of dependent chains of one GEMM shape (graph replay): python tools/chain_gemm.py"""
import ctypes, importlib, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
m = importlib.import_module("mixed-precision_lu_factorization_b200")
lib = m.load_library()
f = lib.mplu_bench_gemm_chain
f.argtypes = [ctypes.c_int] * 9 + [ctypes.POINTER(ctypes.c_float)]
torch.zeros(1, device="cuda")
shapes = [(0, 128, 2048, 128, 1, 0), (0, 128, 30720, 128, 1, 0), (1, 30720, 128, 128, 1, 0), (1, 30720, 128, 128, 0, 0),
          (1, 30720, 256, 256, 1, 1), (1, 30720, 1024, 1024, 1, 1), (1, 1024, 30720, 1024, 1, 1), (0, 128, 128, 128, 1, 0),
          (1, 256, 256, 128, 1, 1), (1, 30720, 2048, 2048, 1, 1), (1, 8192, 128, 128, 1, 0), (1, 8192, 1024, 1024, 1, 1)]
for (v, M, N, K, sh, acc) in shapes:
    for sms in (0, 24):
        out = []
        for pdl in (0, 1):
            us = ctypes.c_float()
            rc = f(v, M, N, K, 50, pdl, sh, acc, sms, ctypes.byref(us))
            assert rc == 0, rc
            out.append(us.value)
        fl = 2.0 * M * N * K
        print(f"variant {v} {M}x{N}x{K} shadow={sh} acc={acc} sms={sms or 148}: pdl0 {out[0]:7.2f} us  pdl1 {out[1]:7.2f} us  "
              f"({fl / out[1] / 1e6:7.1f} TFLOP/s)", flush=True)
