"""CPU tests: the C-ABI library builds, loads and exports every symbol the headers declare (no compute calls)."""
import ctypes
import os
import re
import subprocess

from conftest import ROOT


def _declared_c_functions(header):
    src = open(os.path.join(ROOT, "include", header)).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    src = re.sub(r"//[^\n]*", "", src)
    return re.findall(r"\b(mplu_[a-zA-Z0-9_]+)\s*\(", src)


def test_library_exports_everything_in_mplu_h(mplu):
    lib = mplu.load_library()
    names = set(_declared_c_functions("mplu.h"))
    assert {"mplu_create", "mplu_gesv_device", "mplu_gesv_host", "mplu_gemm16"} <= names
    for name in names:
        assert hasattr(lib, name), f"{name} declared in include/mplu.h but not exported by libmplu.so"


def test_reference_symbols_keep_their_mangled_names(mplu):
    # /root/reference/MPF.h:3, hgetf2_kernel.h:10, dgetf2_native_npv.h:8 have C++ linkage: the link symbols are part
    # of the drop-in contract
    out = subprocess.run(["nm", "-D", "--defined-only", str(mplu.LIB_PATH)], capture_output=True, text=True).stdout
    for sym in ("_Z3MPFPdiiPi", "_Z13HGETF2_kernelP6__halfiiiPi", "_Z17dgetf2_native_npviiPdi"):
        assert sym in out, f"{sym} missing from libmplu.so"


def test_default_options(mplu):
    o = mplu.default_options()
    # nb = 0 means "by n" (resolved at factorization time); pdl is off by default (measured slower)
    assert (o.precision, o.nb, o.max_iters, o.a_exp, o.l_exp) == (0, 0, 30, 11, 11)
    assert (o.lookahead, o.use_graph, o.pdl, o.group) == (1, 1, 0, 1) and o.side_sms % 2 == 0


def test_no_device_is_reported_not_crashed(mplu):
    import torch
    if torch.cuda.is_available():
        return
    lib = mplu.load_library()
    ctx = ctypes.c_void_p()
    assert lib.mplu_create(ctypes.byref(ctx), 0) == -2  # MPLU_E_NODEVICE (the reference prints and returns, MPF.cu:72)


def test_headers_compile_as_c_and_cxx(tmp_path):
    c = tmp_path / "t.c"
    c.write_text('#include "mplu.h"\nint main(void){mplu_options o; mplu_stats s; (void)o; (void)s; return 0;}\n')
    subprocess.run(["gcc", "-std=c99", "-Wall", "-Werror", "-I", os.path.join(ROOT, "include"), "-c", str(c), "-o",
                    str(tmp_path / "t.o")], check=True)
    cpp = tmp_path / "t.cpp"
    cpp.write_text('#include "mplu.h"\n#include "MPF.h"\nint main(){void (*f)(double*,int,int,int*) = &MPF; (void)f; return 0;}\n')
    subprocess.run(["g++", "-std=c++17", "-Wall", "-Werror", "-I", os.path.join(ROOT, "include"), "-c", str(cpp), "-o",
                    str(tmp_path / "t2.o")], check=True)
