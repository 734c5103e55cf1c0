"""Open functor table without a GPU: an objective defined OUTSIDE the library compiles against the public headers alone
(include/pnol/device/functor_kernels.cuh), registers its launch table when its shared library is loaded, and the C-ABI refuses bad
registrations. The reference's plug-in point is subclassing Objective / MultiObjective (Source/PNOL_Objective.hpp:29, :57)."""
import ctypes as C

import numpy as np

import user_functor_lib as U
from parallelnonlinearoptimizationlibrary_b200 import capi


def test_user_objectives_compile_out_of_tree_and_register():
    lib = capi.load_library()
    u = U.build_and_load()
    assert u.my_registration_status() == 0
    for kind in (U.MY_F_TRID, U.MY_F_STYBLINSKI, U.MY_F_GAUSSFIT):
        assert lib.pnol_functor_registered(kind) == 1
    assert lib.pnol_functor_registered(1999) == 0 and lib.pnol_functor_registered(2999) == 0
    assert lib.pnol_functor_registered(capi.F_ROSENBROCK) == 1 and lib.pnol_functor_registered(77) == 0
    # host objEval of the user's functor (the same __host__ __device__ source the kernels compile)
    x = np.array([1.0, 2.0, 3.0])
    assert u.my_host_eval(U.MY_F_TRID, C.c_double(1.0), C.c_void_p(x.ctypes.data), 3) == (0.0 + 1.0 + 4.0) - (2.0 + 6.0)


def test_register_functor_refuses_bad_tables():
    lib = capi.load_library()
    assert lib.pnol_register_functor(1500, None) == capi.ERR_INVALID
    buf = (C.c_char * 256)()                      # abi_version 0, no launchers
    assert lib.pnol_register_functor(1500, C.byref(buf)) == capi.ERR_INVALID
    assert lib.pnol_functor_registered(1500) == 0
