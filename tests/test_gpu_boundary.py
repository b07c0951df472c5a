"""Boundary proofs with the reference's UNMODIFIED sources, built by oracle/Makefile in the build
container (where /root/reference exists) into oracle/_ref/ and run here on the GPU box:

  1. ref_bench_on_mplu        /root/reference/benchmark.cpp linked against libmplu.so: the reference's own driver calls
                              the repo's MPF() (MPF.h:3, benchmark.cpp:220) and runs its own P*L*U == A check on the
                              reference generator's matrices (benchmark.cpp:97-144, 225-231).
  2. ref_mpf_on_mplu_kernels  /root/reference/MPF.cu + benchmark.cpp linked against libmplu_dropin.a: the reference's own
                              host loop launches the repo's HGETF2_kernel / dgetf2_native_npv itself with
                              cudaLaunchCooperativeKernel (MPF.cu:126-133, 178-185) -- a foreign translation unit, the
                              reference's kernels are not compiled in.
"""
import os
import subprocess

import pytest

from conftest import ROOT

pytestmark = pytest.mark.gpu
REF = os.path.join(ROOT, "oracle", "_ref")


def _run_driver(binary, tmp_path, max_size=256):
    exe = os.path.join(REF, binary)
    gen = os.path.join(REF, "matgen")
    if not (os.path.exists(exe) and os.path.exists(gen)):
        pytest.skip(f"oracle/_ref/{binary} not built (make -C oracle where /root/reference exists)")
    mats = str(tmp_path / "mats.txt")
    subprocess.run([gen, mats, str(max_size)], check=True, stdout=subprocess.DEVNULL)  # 2, 4, ..., max_size (step 2, exp)
    r = subprocess.run([exe, mats], cwd=str(tmp_path), capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-2000:]
    rows = open(tmp_path / "benchmark_times.csv").read().strip().splitlines()
    return r.stdout, rows


@pytest.mark.parametrize("binary", ["ref_bench_on_mplu", "ref_mpf_on_mplu_kernels"])
def test_unmodified_reference_sources_against_the_repo(tmp_path, binary):
    out, rows = _run_driver(binary, tmp_path)
    assert rows[0] == "matrix_size,mpf_time,lapack_time"                      # benchmark.cpp:170
    assert [int(l.split(",")[0]) for l in rows[1:]] == [2, 4, 8, 16, 32, 64, 128, 256]
    assert out.count("Checking correctness of MPF results...") == 8           # benchmark.cpp:226
    assert "MPF produced incorrect results." not in out                        # its own 1e-10 check passed every time
    assert "LAPACKE_dgetrf produced incorrect results." not in out
    assert "No CUDA devices available." not in out
