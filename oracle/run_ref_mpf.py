"""Run the UNMODIFIED reference MPF() (oracle/_ref/libmpf_ref.so, built by oracle/Makefile) on one generated input and
save its output.  Test infrastructure only: tests start this in its OWN process because the reference library and
libmplu.so both define the C++ symbols MPF / HGETF2_kernel / dgetf2_native_npv (that is the drop-in contract), and a
process that has libmplu.so loaded RTLD_GLOBAL would interpose them.

    python oracle/run_ref_mpf.py <n> <seed> <out.npz> [r=32] [kind=dd|rand]

kind=dd:   counter_matrix(n, seed, dominant=True)  (same values as mplu_generate on the device)
kind=rand: counter_matrix(n, seed, dominant=False) (pivoting input: exercises the fp16 pivot discovery, MPF.cu:125-163)
Output: LU (fp64, the dgetrf layout MPF returns, MPF.cu:66-256), ipiv (1-based), seconds.
"""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
import mplu_oracle as orc  # noqa: E402
from make_golden import call_ref, load_ref  # noqa: E402


def main():
    n, seed, out = int(sys.argv[1]), int(sys.argv[2]), sys.argv[3]
    kv = dict(a.split("=", 1) for a in sys.argv[4:] if "=" in a)
    r = int(kv.get("r", 32))
    A = orc.counter_matrix(n, seed=seed, dominant=kv.get("kind", "dd") == "dd")
    f = load_ref()
    LU, ipiv, dt = call_ref(f, A, r)
    np.savez(out, LU=LU, ipiv=ipiv, seconds=dt, seed=seed, r=r)


if __name__ == "__main__":
    main()
