// Compact-bounds instantiations of the solve kernel for the padded horizon Tp = 288 (acb_options.path = 3).
#include "acb_solve_kernel.cuh"
ACB_INSTANTIATE_COMPACT_Q(9)
