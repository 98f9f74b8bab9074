// Host side of the batched solve: argument checks, launch configuration, dispatch to the
// per-horizon kernel instantiations (acb_solve_q5.cu, acb_solve_q9.cu), and the standalone
// charging_rate_bounds kernel.
#include <algorithm>
#include "acb_solve_kernel.cuh"

size_t acb_solve_smem_bytes(const SiteDev& s, int Tp, int S_max, int nwarps) {
    return (size_t)make_layout(s.N, s.R, s.NG, s.NP, s.nSlots, Tp, S_max, nwarps).total * sizeof(float);
}

// charging_rate_bounds as a standalone kernel (parity tests; the solve kernel fuses it)
__global__ void acb_bounds_kernel(SiteDev S, acb_batch B, float* lb, float* ub) {
    const int b = blockIdx.x, N = S.N, Tp = B.Tp;
    const int nS = B.n_sessions[b];
    float* lbb = lb + (size_t)b * N * Tp;
    float* ubb = ub + (size_t)b * N * Tp;
    for (int i = threadIdx.x; i < N * Tp; i += blockDim.x) { lbb[i] = 0.f; ubb[i] = 0.f; }
    __syncthreads();
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, nwarps = blockDim.x >> 5;
    for (int s = warp; s < nS; s += nwarps) {
        size_t k = (size_t)b * B.S_max + s;
        int row = B.sess_row[k], a = B.sess_start[k], len = B.sess_len[k], off = B.sess_rate_off[k];
        for (int j = lane; j < len; j += 32) {
            const int ri = off >= 0 ? off + j : -(off + 1);  // off < 0: one (min, max) pair for the whole session
            float lo = B.min_rates[ri], hi = B.max_rates[ri];
            if (a + j < Tp) { lbb[row * Tp + a + j] = lo; ubb[row * Tp + a + j] = fmaxf(hi, lo); }
        }
    }
}

int acb_launch_solve_q2(const acb_site*, const acb_batch*, const acb_options*, int, size_t, cudaStream_t, bool, int);
int acb_launch_solve_q4(const acb_site*, const acb_batch*, const acb_options*, int, size_t, cudaStream_t, bool, int);
int acb_launch_solve_q5(const acb_site*, const acb_batch*, const acb_options*, int, size_t, cudaStream_t, bool, int);
int acb_launch_solve_compact_q4(const acb_site*, const acb_batch*, const acb_options*, int, size_t, cudaStream_t, int);
int acb_launch_solve_compact_q9(const acb_site*, const acb_batch*, const acb_options*, int, size_t, cudaStream_t, int);
int acb_launch_solve_q9(const acb_site*, const acb_batch*, const acb_options*, int, size_t, cudaStream_t, bool, int);

extern "C" int acb_solve_batch(acb_site* site, const acb_batch* batch, const acb_options* opt_in, void* stream) {
    if (!site || !batch || batch->B <= 0 || batch->Tp <= 0 || batch->Tp % 32 != 0) {
        acb_set_error("acb_solve_batch: bad arguments (Tp must be a positive multiple of 32)");
        return ACB_E_INVALID;
    }
    if ((site->d.has_pl && !batch->peak_limit)) {
        acb_set_error("acb_solve_batch: site was created with use_peak_row but batch.peak_limit is NULL");
        return ACB_E_INVALID;
    }
    acb_options opt;
    if (opt_in) opt = *opt_in; else acb_default_options(&opt);
    ACB_CUDA(cudaSetDevice(site->device));
    const SiteDev& d = site->d;
    const int nCT_ = d.nDisc + d.nLin + d.has_pl + d.has_u;
    const int nCT = nCT_;
    const int Q = batch->Tp / 32;
    if (batch->Tp % 32 != 0 || (Q != 2 && Q != 4 && Q != 5 && Q != 9)) {
        acb_set_error("acb_solve_batch: Tp must be 64, 128, 160 or 288 (pad the horizon up)");
        return ACB_E_INVALID;
    }
    cudaStream_t st = (cudaStream_t)stream;
    const int NIN = d.NG + d.R;
    const int nch = (NIN <= ACB_OPP) ? 1 : 3;
    const int nParts = (NIN + ACB_OPP * nch - 1) / (ACB_OPP * nch);
    // threads: warps for the EVSE rows plus room for the coupling rows, and one column-pass sweep if possible
    int want = std::max(d.nRowWarps * 32 + nCT * 16, std::min(1024, nParts * batch->Tp));
    int nthreads = std::min(768, ((want + 31) / 32) * 32);
    if (opt.path == 3) {
        // experimental compact-bounds kernel: no materialised lb/ub (constant limits, one session per EVSE), 6 rows per
        // warp, 384 threads, two blocks per SM.  Opt-in only; the caller guarantees the batch qualifies.
        if (!site->has_d6 || batch->multi_session || (Q != 4 && Q != 9)) {
            acb_set_error("acb_solve_batch: path 3 (compact bounds) needs a site of <= 192 EVSEs, one session per EVSE and Tp 128 or 288");
            return ACB_E_INVALID;
        }
        const SiteDev& e = site->d6;
        const int nt = 384;
        const size_t sm6 = (size_t)make_layout(e.N, e.R, e.NG, e.NP, e.nSlots, batch->Tp, batch->S_max, nt / 32, true).total * sizeof(float);
        if (e.nRowWarps * 32 > nt || nCT_ > 32 || sm6 > 113 * 1024) {
            acb_set_error("acb_solve_batch: path 3 (compact bounds) does not fit two blocks per SM for this site (" + std::to_string(sm6) + " B)");
            return ACB_E_TOO_LARGE;
        }
        return (Q == 4) ? acb_launch_solve_compact_q4(site, batch, &opt, nt, sm6, st, nch) : acb_launch_solve_compact_q9(site, batch, &opt, nt, sm6, st, nch);
    }
    size_t smem = acb_solve_smem_bytes(d, batch->Tp, batch->S_max, nthreads / 32);
    const bool fits = d.TPW == 3 && nCT_ <= 32 && nthreads <= 768 && d.nRowWarps * 32 <= nthreads && smem <= 232448;
    if (opt.path == 2 || (!fits && opt.path == 0)) return acb_solve_general(site, batch, opt, st);
    if (!fits) {
        acb_set_error("acb_solve_batch: instance does not fit the on-chip path (N <= ~66 EVSEs, <= 32 coupling tasks, " +
                      std::to_string(smem) + " B of shared memory needed, 232448 available)");
        return ACB_E_TOO_LARGE;
    }
    const bool multi = batch->multi_session != 0;
    if (Q == 2) return acb_launch_solve_q2(site, batch, &opt, nthreads, smem, st, multi, nch);
    if (Q == 4) return acb_launch_solve_q4(site, batch, &opt, nthreads, smem, st, multi, nch);
    if (Q == 5) return acb_launch_solve_q5(site, batch, &opt, nthreads, smem, st, multi, nch);
    return acb_launch_solve_q9(site, batch, &opt, nthreads, smem, st, multi, nch);
}

extern "C" int acb_charging_rate_bounds(acb_site* site, const acb_batch* batch, float* lb, float* ub, void* stream) {
    if (!site || !batch || !lb || !ub) { acb_set_error("acb_charging_rate_bounds: bad arguments"); return ACB_E_INVALID; }
    ACB_CUDA(cudaSetDevice(site->device));
    acb_bounds_kernel<<<batch->B, 256, 0, (cudaStream_t)stream>>>(site->d, *batch, lb, ub);
    ACB_CUDA(cudaGetLastError());
    return ACB_OK;
}
