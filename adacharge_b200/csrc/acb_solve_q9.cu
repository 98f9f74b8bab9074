// Instantiations of the solve kernel for the padded horizon Tp = 288.
#include "acb_solve_kernel.cuh"
ACB_INSTANTIATE_Q(9)

#ifdef ACB_TRACE
// development build (make trace): clock stamps of block 0, see ACB_TR in acb_solve_kernel.cuh
extern "C" int acb_trace_fetch(long long* out, int n) {
    const int have = ACB_TR_NIT * 32 * ACB_TR_SLOTS;
    cudaDeviceSynchronize();
    return cudaMemcpyFromSymbol(out, g_acb_trace, sizeof(long long) * (n < have ? n : have)) == cudaSuccess ? (n < have ? n : have) : -1;
}
#endif
