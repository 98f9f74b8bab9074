// Instantiations of the solve kernel for the padded horizon Tp = 64.
#include "acb_solve_kernel.cuh"
ACB_INSTANTIATE_Q(2)
