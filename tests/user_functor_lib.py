"""Builds tests/user_functor/my_objectives.cu OUT OF TREE (a temporary directory, the user's own nvcc command line) and loads it:
objectives that are not built into libpnol_b200.so register their launch tables at load time (include/pnol_b200.h:
pnol_register_functor; include/pnol/device/functor_kernels.cuh)."""
import ctypes as C
import os
import shutil
import subprocess
import tempfile

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
SRC = os.path.join(ROOT, "tests", "user_functor")
LIBDIR = os.path.join(ROOT, "parallelnonlinearoptimizationlibrary_b200", "lib")
MY_F_TRID, MY_F_STYBLINSKI, MY_F_GAUSSFIT = 1001, 1002, 2001

_lib = None
_dir = None


def nvcc():
    return shutil.which("nvcc") or "/usr/local/cuda/bin/nvcc"


def build_and_load():
    """compile once per process; returns the ctypes handle of the user's library"""
    global _lib, _dir
    if _lib is not None:
        return _lib
    from parallelnonlinearoptimizationlibrary_b200 import capi, hostapi
    capi.load_library()
    hostapi.lib()
    _dir = tempfile.mkdtemp(prefix="pnol_user_functor_")
    out = os.path.join(_dir, "libmy_objectives.so")
    cmd = [nvcc(), "-gencode", "arch=compute_100a,code=sm_100a", "-std=c++17", "-fmad=false", "-Xcompiler", "-fPIC,-ffp-contract=off",
           "-I" + os.path.join(ROOT, "include"), "-I" + SRC, "-shared", "-o", out, os.path.join(SRC, "my_objectives.cu"),
           "-L" + LIBDIR, "-lpnol_b200", "-lpnol_b200_host", "-Xlinker", "-rpath", "-Xlinker", LIBDIR]
    subprocess.run(cmd, check=True, cwd=_dir, capture_output=True, text=True)
    _lib = C.CDLL(out, mode=C.RTLD_GLOBAL)
    _lib.my_host_eval.restype = C.c_double
    return _lib
