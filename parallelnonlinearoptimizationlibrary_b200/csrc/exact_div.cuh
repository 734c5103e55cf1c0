// moved to the public device headers (a user functor's out-of-tree kernels need it: include/pnol/device/functor_kernels.cuh)
#pragma once
#include "pnol/device/exact_div.cuh"
