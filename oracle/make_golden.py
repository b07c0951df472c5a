"""Generate tests/golden/*.npz by running the UNMODIFIED reference (oracle/_ref/libmpf_ref.so, built from
/root/reference by oracle/Makefile) on a B200.  Test infrastructure only.

    gpurun -- python oracle/make_golden.py        # writes gpurun_out/golden/*.npz ; copy them to tests/golden/

Fixtures (inputs are regenerated from seeds by the oracle, so only outputs are stored):
  ref_mpf_dd_n{128,256}.npz    MPF(A, n, 32, ipiv) on counter_matrix(n, seed=1, dominant=True): LU (fp64), ipiv
  ref_mpf_rand_n{64,128,256}.npz   MPF on the reference generator's own stream (`matrix_generator f 256 2 exp`,
                               read the way benchmark.cpp reads it): LU, ipiv  -- exercises fp16 pivot discovery
  ref_timing.json              wall time of the reference's MPF() per size (benchmark.cpp:219-222 semantics)
"""
import ctypes
import json
import os
import sys
import time

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
sys.path.insert(0, HERE)
import mplu_oracle as orc  # noqa: E402


def load_ref():
    lib = ctypes.CDLL(os.path.join(HERE, "_ref", "libmpf_ref.so"))
    f = getattr(lib, "_Z3MPFPdiiPi")  # void MPF(double*, int, int, int*)  (MPF.h:3, C++ linkage)
    f.argtypes = [ctypes.c_void_p, ctypes.c_int, ctypes.c_int, ctypes.c_void_p]
    f.restype = None
    return f


def call_ref(f, A, r=32):
    n = A.shape[0]
    Af = np.asfortranarray(A.copy())
    ipiv = np.arange(1, n + 1, dtype=np.int32)  # benchmark.cpp:215-217
    sys.stdout.flush()
    devnull = os.open(os.devnull, os.O_WRONLY)
    saved = os.dup(1)
    os.dup2(devnull, 1)  # MPF prints one line per panel (MPF.cu:137)
    t = time.perf_counter()
    f(Af.ctypes.data, n, r, ipiv.ctypes.data)
    dt = time.perf_counter() - t
    os.dup2(saved, 1)
    os.close(saved)
    os.close(devnull)
    return Af, ipiv, dt


def main():
    out = os.path.join(ROOT, "gpurun_out", "golden")
    os.makedirs(out, exist_ok=True)
    f = load_ref()
    call_ref(f, orc.counter_matrix(64, 1))  # warm-up: context + cuBLAS init
    timing = {}
    for n in (128, 256):
        A = orc.counter_matrix(n, seed=1)
        LU, ipiv, dt = call_ref(f, A)
        o_lu, o_ip = orc.mpf_reference(A, 32)
        print(f"dd n={n}: identity pivots {bool((ipiv == np.arange(1, n + 1)).all())} |ref-oracle| {np.abs(LU - o_lu).max():.2e} "
              f"check {orc.check_correctitude(A, LU, ipiv)}")
        np.savez_compressed(os.path.join(out, f"ref_mpf_dd_n{n}.npz"), LU=LU, ipiv=ipiv, seed=1, r=32)
    for n, M in orc.matrix_generator_stream(256):
        if n < 64:
            continue
        A = orc.as_benchmark_reads(M)
        LU, ipiv, dt = call_ref(f, A)
        o_lu, o_ip = orc.mpf_reference(A, 32)
        print(f"rand n={n}: check {orc.check_correctitude(A, LU, ipiv)} ipiv==oracle {bool((ipiv == o_ip).all())} "
              f"|ref-oracle| {np.abs(LU - o_lu).max():.2e}")
        np.savez_compressed(os.path.join(out, f"ref_mpf_rand_n{n}.npz"), LU=LU, ipiv=ipiv, r=32)
    sizes = [int(s) for s in sys.argv[1:]] or [1024, 4096, 8192]
    for n in sizes:
        A = orc.counter_matrix(n, seed=1)
        LU, ipiv, dt = call_ref(f, A)
        nonid = int((ipiv != np.arange(1, n + 1)).sum())
        timing[str(n)] = dict(seconds=dt, tflops=2 / 3 * n ** 3 / dt / 1e12, non_identity_pivots=nonid)
        print(f"timing n={n}: {dt:.3f} s  {timing[str(n)]['tflops']:.3f} TFLOP/s  non-identity pivots {nonid}", flush=True)
    json.dump(timing, open(os.path.join(out, "ref_timing.json"), "w"), indent=1)


if __name__ == "__main__":
    main()
