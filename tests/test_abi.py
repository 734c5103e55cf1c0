"""The C-ABI boundary without a GPU: libpnol_b200.so loads and exports every symbol include/pnol_b200.h declares, the
host-class library exports its C face, and compute entry points fail loudly (no CPU fallback) when no device exists."""
import ctypes as C
import os
import re

import numpy as np
import pytest

from parallelnonlinearoptimizationlibrary_b200 import capi

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
HEADER = os.path.join(ROOT, "include", "pnol_b200.h")

HOST_EXPORTS = ["pnolhost_last_error", "pnolhost_attach", "pnolhost_detach", "pnolhost_set_pool_width", "pnolhost_set_hinv_mode",
                "pnolhost_set_jac_mode", "pnolhost_set_jacobian_cache", "pnolhost_set_store_jacobian", "pnolhost_set_stream", "pnolhost_lm_lorentz", "pnolhost_lm_problem_create", "pnolhost_lm_problem_run", "pnolhost_lm_problem_destroy",
                "pnolhost_lm_example",
                "pnolhost_gradient", "pnolhost_gradient_recur", "pnolhost_hessian", "pnolhost_obj_eval", "pnolhost_jacobian_example",
                "pnolhost_bfgs", "pnolhost_ga", "pnolhost_simplex", "pnolhost_check_box_bounds", "pnolhost_compute_alpha_bnd"]


def header_symbols():
    src = open(HEADER).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(pnol_[a-z0-9_]+)\s*\(", src)))


def test_header_and_binding_agree():
    assert header_symbols() == sorted(capi.EXPORTS)


def test_library_exports_every_declared_symbol():
    lib = capi.load_library()
    for name in header_symbols():
        assert hasattr(lib, name), "libpnol_b200.so does not export %s" % name


def test_host_library_exports():
    from parallelnonlinearoptimizationlibrary_b200 import hostapi
    h = hostapi.lib()
    for name in HOST_EXPORTS:
        assert hasattr(h, name), "libpnol_b200_host.so does not export %s" % name


def test_signatures_are_plain_c():
    # no C++ / torch types cross the boundary: only scalars, pointers to scalars and the opaque handles
    src = re.sub(r"/\*.*?\*/", "", open(HEADER).read(), flags=re.S)
    assert "std::" not in src and "torch" not in src and "at::" not in src
    assert 'extern "C"' in src


def test_version_and_host_helpers_work_without_gpu():
    lib = capi.load_library()
    assert lib.pnol_version().decode().startswith("pnol_b200")
    # the box helpers and the counter stream are host arithmetic and need no device
    x = np.array([0.0, 5.0, -7.0])
    lb, ub = np.full(3, -1.0), np.full(3, 1.0)
    nrep = C.c_int()
    assert lib.pnol_check_box_bounds(C.c_void_p(x.ctypes.data), C.c_void_p(lb.ctypes.data), C.c_void_p(ub.ctypes.data), 3, C.byref(nrep)) == 0
    assert nrep.value == 2 and np.array_equal(x, [0.0, 0.0, 0.0])
    u = lib.pnol_stream_uniform(C.c_uint64(12345), C.c_uint64(7), C.c_double(0.999))
    assert 0.0 <= u < 0.999


@pytest.mark.skipif(os.path.exists("/dev/nvidiactl"), reason="a GPU is present")
def test_no_cpu_fallback():
    with pytest.raises(capi.PnolError):
        capi.Context(0)
    from parallelnonlinearoptimizationlibrary_b200 import hostapi
    hostapi.detach()
    with pytest.raises(capi.PnolError):
        hostapi.gradient("rosenbrock", np.ones(4), np.full(4, 1e-6))


def test_syrk_stream_k_plan_invariants():
    # host logic of the J^T J kernel's work distribution (dmma.cu: syrk_streamk_plan), checked without a device: every K chunk of
    # every tile role exactly once and in order, contiguous slots per role, at most 8 segments per CTA, shares within one chunk
    lib = capi.load_library()
    lib.pnol_selftest_syrk_plan.restype = C.c_int
    lib.pnol_selftest_syrk_plan.argtypes = [C.c_longlong, C.c_int, C.c_int, C.c_int]
    rng = np.random.default_rng(5)
    shapes = [(1, 16), (31, 16), (32, 16), (33, 128), (4737, 256), (100_000, 48), (7000, 272), (5000, 512), (300_000, 256),
              (4_000_000, 256), (500_000, 256), (4_000_000, 16), (1_000_003, 640), (2_000_000_000, 32), (100_000, 4096),
              (50_000, 8192), (777, 16384)]
    shapes += [(int(rng.integers(1, 3_000_000)), int(16 * rng.integers(1, 45))) for _ in range(200)]
    for m, n in shapes:
        for sms in (1, 2, 7, 132, 148):
            for with_f in (0, 1):
                assert lib.pnol_selftest_syrk_plan(m, n, sms, with_f) == 0, (m, n, sms, with_f)
    assert lib.pnol_selftest_syrk_plan(0, 16, 148, 0) == -1


def test_reference_examples_compile_against_the_plugin_headers_and_fail_loudly_without_a_gpu():
    # oracle/_ref/pnol_examples_dropin = the reference's own Source/Examples.cpp, unmodified, compiled against include/pnol and linked
    # to the two product libraries (oracle/dropin_examples.cpp; `make -C oracle dropin` where /root/reference exists). That it exists
    # says the plugin API is source-compatible with every class and helper the reference's drivers use; without a device a driver
    # that needs the hot path must stop with the library's error, not fall back to the CPU.
    import subprocess
    import torch
    exe = os.path.join(ROOT, "oracle", "_ref", "pnol_examples_dropin")
    if os.path.isdir("/root/reference/Source"):
        subprocess.check_call(["make", "-s", "-C", os.path.join(ROOT, "oracle"), "dropin"])
    if not os.path.exists(exe):
        pytest.skip("oracle/_ref/pnol_examples_dropin not built (needs /root/reference)")
    if torch.cuda.is_available():
        pytest.skip("a GPU is present: the drivers run (tests/test_gpu_dropin_examples.py)")
    for driver in ("testBFGS", "testLMExpMPI", "testGA", "testBFGSBnd_MPI", "testSimplexSearch", "testGradientEvaluation"):
        r = subprocess.run([exe, driver], capture_output=True, text=True, timeout=60, cwd=ROOT)
        assert r.returncode == 1 and "no CPU fallback" in r.stderr, (driver, r.returncode, r.stderr[-300:])
    r = subprocess.run([exe, "noSuchDriver"], capture_output=True, text=True, timeout=60, cwd=ROOT)
    assert r.returncode == 2
