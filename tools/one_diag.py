"""diag_lu launches (128x128 diagonal-block LU + inverses) with per-phase clock64 stamps: python tools/one_diag.py [reps]"""
import ctypes, importlib, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
m = importlib.import_module("mixed-precision_lu_factorization_b200")
lib = m.load_library()
f = lib.mplu_diag_lu128_timed
f.argtypes = [ctypes.c_void_p, ctypes.c_longlong] + [ctypes.c_void_p] * 4
reps = int(sys.argv[1]) if len(sys.argv) > 1 else 3
evict = (sys.argv[2] != "warm") if len(sys.argv) > 2 else True
torch.manual_seed(0)
A = (torch.rand(128, 128, device="cuda") * 9.9 + torch.eye(128, device="cuda") * 700).t().contiguous().t()
Li = torch.zeros(128, 128, device="cuda").t(); Ui = torch.zeros(128, 128, device="cuda").t()
junk = torch.empty(64 << 20, device="cuda")
for _ in range(reps):
    W = A.clone()
    clk = torch.zeros(32, dtype=torch.int64, device="cuda")
    if evict: junk.normal_()  # evict caches (incl. instruction lines) like the interleaved GEMMs do
    torch.cuda.synchronize()
    assert f(W.data_ptr(), W.stride(1), Li.data_ptr(), Ui.data_ptr(), clk.data_ptr(), None) == 0
    torch.cuda.synchronize()
    c = clk.cpu().tolist()
    c = [x for x in c if x]
    print("phase cycles:", [c[i + 1] - c[i] for i in range(len(c) - 1)], "total", c[-1] - c[0])
