"""Exploratory GPU probe (development aid, not part of the test-suite): python tools/probe.py <what> [args]"""
import ctypes, importlib, os, sys, time
import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "oracle"))
m = importlib.import_module("mixed-precision_lu_factorization_b200")
import mplu_oracle as orc

dev = "cuda"


def cm(t):  # column-major copy of a 2-D tensor (same logical values)
    return t.t().contiguous().t()


def gemm():
    torch.manual_seed(0)
    for variant in (0, 1, 2, 3):
        for (M, N, K) in [(128, 256, 64), (256, 256, 128), (384, 512, 256), (1000, 700, 192), (4096, 4096, 1024)]:
            for dt in (torch.float16, torch.bfloat16):
                A = cm(torch.randn(M, K, device=dev).to(dt))
                B = cm(torch.randn(K, N, device=dev).to(dt))
                C0 = cm(torch.randn(M, N, device=dev))
                C = C0.clone()
                C = cm(C)
                try:
                    if variant in (2, 3):
                        At = cm(A.t().contiguous())  # K x M column-major
                        out, H = m.gemm16(variant, None, B, C, alpha=-0.5, beta=1.0, want_shadow=True, hscale=0.25, a_transposed=At)
                    else:
                        out, H = m.gemm16(variant, A, B, C, alpha=-0.5, beta=1.0, want_shadow=True, hscale=0.25)
                    ref = C0.double() - 0.5 * (A.double() @ B.double())
                    err = (out.double() - ref).abs().max().item() / ref.abs().max().item()
                    herr = (H.double() - 0.25 * ref).abs().max().item() / ref.abs().max().item()
                    print(f"variant {variant} {M}x{N}x{K} {str(dt)[6:]}: rel err {err:.2e} shadow err {herr:.2e}", flush=True)
                except Exception as e:
                    print(f"variant {variant} {M}x{N}x{K} {dt}: EXC {e}", flush=True)


def gemm_perf():
    for variant in (0, 1):
        for (M, N, K) in [(8192, 8192, 1024), (16384, 16384, 1024), (16384, 16384, 2048), (30720, 30720, 1024), (30720,30720,2048), (32640, 128, 128), (128, 30000, 128)]:
            A = cm(torch.randn(M, K, device=dev).half())
            B = cm(torch.randn(K, N, device=dev).half())
            C = cm(torch.zeros(M, N, device=dev))
            for sh in (False, True):
                m.gemm16(variant, A, B, C, alpha=-1.0, beta=1.0, want_shadow=False)
                lib = m.load_library()
                H = cm(torch.zeros(M, N, device=dev).half()) if sh else None
                torch.cuda.synchronize()
                e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                st = torch.cuda.current_stream().cuda_stream
                e0.record()
                reps = 5
                for _ in range(reps):
                    lib.mplu_gemm16(variant, 0, M, N, K, -1.0, A.data_ptr(), A.stride(1), B.data_ptr(), B.stride(1), 1.0,
                                    C.data_ptr(), C.stride(1), H.data_ptr() if sh else None, H.stride(1) if sh else 0, 1.0, 0, st)
                e1.record()
                torch.cuda.synchronize()
                ms = e0.elapsed_time(e1) / reps
                print(f"perf variant {variant} {M}x{N}x{K} shadow={sh}: {ms:.3f} ms  {2*M*N*K/ms/1e9:.1f} TFLOP/s  C-traffic {8*M*N/ms/1e6:.0f} GB/s", flush=True)
            del A, B, C


def diag():
    lib = m.load_library()
    A = orc.counter_matrix(128, seed=5)
    W = torch.tensor(A, dtype=torch.float32, device=dev).t().contiguous().t()
    Li = torch.zeros(128, 128, device=dev).t(); Ui = torch.zeros(128, 128, device=dev).t()
    Li = cm(Li); Ui = cm(Ui)
    torch.cuda.synchronize()
    rc = lib.mplu_diag_lu128(W.data_ptr(), W.stride(1), Li.data_ptr(), Ui.data_ptr(), None)
    torch.cuda.synchronize()
    ref = orc.dgetf2_npv(A.astype(np.float32).astype(np.float64))
    got = W.cpu().double().numpy()
    print("diag rc", rc, "LU rel err", np.abs(got - ref).max() / np.abs(ref).max())
    L = np.tril(ref, -1) + np.eye(128); U = np.triu(ref)
    print("Linv err", np.abs(Li.cpu().double().numpy() - np.linalg.inv(L)).max(), "Uinv err", np.abs(Ui.cpu().double().numpy() - np.linalg.inv(U)).max() / np.abs(np.linalg.inv(U)).max())
    # random non-dominant block too
    R = np.random.default_rng(0).standard_normal((128, 128)) + 12 * np.eye(128)
    W = cm(torch.tensor(R, dtype=torch.float32, device=dev))
    lib.mplu_diag_lu128(W.data_ptr(), W.stride(1), Li.data_ptr(), Ui.data_ptr(), None)
    torch.cuda.synchronize()
    ref = orc.dgetf2_npv(R.astype(np.float32).astype(np.float64))
    print("diag(random) LU rel err", np.abs(W.cpu().double().numpy() - ref).max() / np.abs(ref).max())


def gesv(ns=(256, 1024, 4096), nb=1024, variant=-1, prec=0):
    s = m.Solver(0)
    for n in ns:
        A = torch.tensor(orc.counter_matrix(n, seed=1), dtype=torch.float64, device=dev)
        Acm = cm(A)
        b = A.sum(dim=1)
        opts = m.default_options(nb=nb, gemm_variant=variant, precision=prec)
        try:
            x, st = s.gesv(Acm, b, opts, allow_noconv=True)
            d = st.as_dict()
            print(f"gesv n={n} nb={nb} var={variant} prec={prec}: iters {d['iters']} conv {d['converged']} be {d['backward_error']:.2e} first {d['first_backward_error']:.2e} "
                  f"status {d['status_bits']} factor {d['factor_ms']:.2f} ms solve {d['solve_ms']:.2f} ms  TF {2/3*n**3/d['total_ms']/1e9:.2f} (factor only {2/3*n**3/d['factor_ms']/1e9:.2f}) launches {d['kernel_launches']} fwd err {(x-1).abs().max().item():.2e}", flush=True)
            if n <= 4096:
                LU = s.factors(n).cpu().numpy()
                ref = orc.lu_nopivot_fp64(A.cpu().numpy())
                print(f"   factor err vs fp64 no-pivot LU: max abs {np.abs(LU-ref).max():.3e} rel(max) {np.abs(LU-ref).max()/np.abs(ref).max():.3e}  L part {np.abs(np.tril(LU-ref,-1)).max():.3e}/{np.abs(np.tril(ref,-1)).max():.3e}", flush=True)
        except Exception as e:
            print(f"gesv n={n}: EXC {e}", flush=True)
        del A, Acm


def gesv_dev(ns=(8192,), nb=1024, variant=-1, prec=0, reps=2):
    s = m.Solver(0)
    for n in ns:
        A, b = m.generate(n, seed=1)
        opts = m.default_options(nb=nb, gemm_variant=variant, precision=prec)
        for rep in range(reps):
            try:
                x, st = s.gesv(A, b, opts, allow_noconv=True)
                d = st.as_dict()
                print(f"gesv n={n} nb={nb} var={variant} prec={prec} rep={rep}: iters {d['iters']} conv {d['converged']} be {d['backward_error']:.2e} first {d['first_backward_error']:.2e} "
                      f"status {d['status_bits']} factor {d['factor_ms']:.2f} ms solve {d['solve_ms']:.2f} ms  TF {2/3*n**3/d['total_ms']/1e9:.2f} (factor only {2/3*n**3/d['factor_ms']/1e9:.2f}) launches {d['kernel_launches']} gemms {d['gemm_launches']} fwd err {(x-1).abs().max().item():.2e}", flush=True)
            except Exception as e:
                print(f"gesv n={n}: EXC {e}", flush=True)
        del A, b


def ref(ns=(256, 1024, 4096)):
    lib = ctypes.CDLL(os.path.join(ROOT, "oracle", "_ref", "libmpf_ref.so"))
    f = getattr(lib, "_Z3MPFPdiiPi")
    f.argtypes = [ctypes.c_void_p, ctypes.c_int, ctypes.c_int, ctypes.c_void_p]
    f.restype = None
    devnull = os.open(os.devnull, os.O_WRONLY)
    for n in ns:
        A = orc.counter_matrix(n, seed=1)
        Af = np.asfortranarray(A.copy())
        ipiv = np.arange(1, n + 1, dtype=np.int32)
        sys.stdout.flush()
        saved = os.dup(1); os.dup2(devnull, 1)
        t = time.time()
        f(Af.ctypes.data, n, 32, ipiv.ctypes.data)
        dt = time.time() - t
        os.dup2(saved, 1); os.close(saved)
        nonid = int((ipiv != np.arange(1, n + 1)).sum())
        line = f"ref MPF n={n}: {dt:.3f} s ({2/3*n**3/dt/1e12:.3f} TFLOP/s) non-identity pivots {nonid}"
        if n <= 2048:
            LU, ip = orc.mpf_reference(A, 32)
            line += f" | vs oracle: max abs diff {np.abs(Af-LU).max():.3e} ipiv equal {bool((ip==ipiv).all())} check {orc.check_correctitude(A, Af, ipiv)}"
        print(line, flush=True)
        if n in (128, 256):
            np.savez_compressed(os.path.join(ROOT, "gpurun_out", f"ref_mpf_dd_n{n}.npz"), LU=Af, ipiv=ipiv, seed=1)
    # non-dominant (pivoting) case from the reference's own generator stream
    for n, M in orc.matrix_generator_stream(256):
        if n < 64: continue
        A = orc.as_benchmark_reads(M)
        Af = np.asfortranarray(A.copy()); ipiv = np.arange(1, n + 1, dtype=np.int32)
        sys.stdout.flush(); saved = os.dup(1); os.dup2(devnull, 1)
        f(Af.ctypes.data, n, 32, ipiv.ctypes.data)
        os.dup2(saved, 1); os.close(saved)
        LU, ip = orc.mpf_reference(A, 32, fused=True)
        LU2, ip2 = orc.mpf_reference(A, 32, fused=False)
        print(f"ref MPF rand n={n}: check {orc.check_correctitude(A, Af, ipiv)} ipiv==oracle(fused) {bool((ip==ipiv).all())} ({int((ip!=ipiv).sum())} differ) ipiv==oracle(unfused) {bool((ip2==ipiv).all())} max diff(fused) {np.abs(Af-LU).max():.3e}", flush=True)
        np.savez_compressed(os.path.join(ROOT, "gpurun_out", f"ref_mpf_rand_n{n}.npz"), LU=Af, ipiv=ipiv)


if __name__ == "__main__":
    os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
    what = sys.argv[1]
    args = [eval(a) for a in sys.argv[2:]]
    print(f"== {what} {args}", flush=True)
    globals()[what](*args)
