"""B200-native (sm_100a) implementation of PNOL's data-parallel evaluation-and-derivative hot path.

The product is C++/CUDA: `lib/libpnol_b200.so` (hand-written kernels behind the C-ABI of include/pnol_b200.h) and
`lib/libpnol_b200_host.so` (the host C++ mirror of the reference's PNOL_Objective / PNOL_Algorithm plugin API).
This Python package only binds them (ctypes) for tests, the benchmark and torch.distributed launch plumbing."""
from . import capi, problems  # noqa: F401

__all__ = ["capi", "problems"]
