"""Isolated rate of the tall rank-nb update for the current MPLU_GEMM_PAIRS_PER_CLUSTER: python tools/gemm_cluster_probe.py"""
import ctypes, importlib, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
m = importlib.import_module("mixed-precision_lu_factorization_b200")
lib = m.load_library()
f = lib.mplu_bench_gemm_chain
f.argtypes = [ctypes.c_int] * 9 + [ctypes.POINTER(ctypes.c_float)]
torch.zeros(1, device="cuda")
for (v, M, N, K, sh, acc) in [(1, 30720, 2048, 2048, 0, 1), (1, 30720, 6144, 2048, 0, 1), (1, 30720, 2048, 2048, 1, 1), (1, 16384, 4096, 2048, 0, 1)]:
    for sms in (0, 132):
        us = ctypes.c_float()
        best = 1e30
        for _ in range(3):
            rc = f(v, M, N, K, 20, 0, sh, acc, sms, ctypes.byref(us))
            assert rc == 0, rc
            best = min(best, us.value)
        print(f"pairs/cluster={os.environ.get('MPLU_GEMM_PAIRS_PER_CLUSTER', '1')} {M}x{N}x{K} shadow={sh} sms={sms or 148}: {best:8.2f} us  {2.0 * M * N * K / best / 1e6:7.1f} TFLOP/s", flush=True)
