// Instantiations of the solve kernel for the padded horizon Tp = 160.
#include "acb_solve_kernel.cuh"
ACB_INSTANTIATE_Q(5)
