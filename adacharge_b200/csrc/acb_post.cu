// Postprocessing kernels, float64 like the reference (numpy defaults), written so that
// the results are bit-identical to reference adacharge/postprocessing.py given the same
// continuous schedule:
//   acb_project_continuous  <- project_into_continuous_feasible_pilots  pp.py:77-94
//   acb_project_discrete    <- project_into_discrete_feasible_pilots    pp.py:97-118
//                              (floor_to_set pp.py:10-31, eps = 0.05)
//   acb_reallocate          <- index_based_reallocation pp.py:121-186 (mode 0),
//                              diff_based_reallocation  pp.py:189-258 (mode 1)
//                              (increment_in_set pp.py:58-74)
//   acb_constraints_feasible<- infrastructure_constraints_feasible      utils.py:5-12
// Arithmetic notes: sums that the reference takes with np.sum follow numpy's pairwise
// summation order; products/sums are issued without FMA contraction.
#include <algorithm>
#include "acb_common.cuh"

__device__ __forceinline__ int bisect_left(const double* a, int n, double x) {
    int lo = 0, hi = n;
    while (lo < hi) { int mid = (lo + hi) >> 1; if (a[mid] < x) lo = mid + 1; else hi = mid; }
    return lo;
}
__device__ __forceinline__ int bisect_right(const double* a, int n, double x) {
    int lo = 0, hi = n;
    while (lo < hi) { int mid = (lo + hi) >> 1; if (x < a[mid]) hi = mid; else lo = mid + 1; }
    return lo;
}
__device__ __forceinline__ double floor_to_set(double x, const double* set, int n, double eps) {
    int pos = bisect_left(set, n, __dadd_rn(x, eps));
    if (pos < n && x == set[pos]) return x;
    if (pos == 0) return set[0];
    if (pos == n) return set[n - 1];
    return set[pos - 1];
}
__device__ __forceinline__ double increment_in_set(double x, const double* set, int n) {
    int pos = bisect_right(set, n, x);
    if (pos == 0) return set[0];
    if (pos == n) return set[n - 1];
    return set[pos];
}

__global__ void k_project_continuous(SiteDev S, const double* in, double* out, int B, int T) {
    size_t total = (size_t)B * S.N * T;
    for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < total; i += (size_t)gridDim.x * blockDim.x) {
        int row = (int)((i / T) % S.N);
        double v = in[i];
        double mp = S.max_pilot[row];
        v = (mp < v) ? mp : v;          // np.minimum
        out[i] = (v > 0.0) ? v : 0.0;   // np.maximum(., 0)
    }
}

__global__ void k_project_discrete(SiteDev S, const double* in, double* out, int B, int T) {
    size_t total = (size_t)B * S.N * T;
    for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < total; i += (size_t)gridDim.x * blockDim.x) {
        int row = (int)((i / T) % S.N);
        int o = S.allow_off[row], n = S.allow_off[row + 1] - o;
        double v = floor_to_set(in[i], S.allow_vals + o, n, 0.05);
        out[i] = (v > 0.0) ? v : 0.0;
    }
}

// numpy's pairwise sum of n float64 values (numpy/core/src/umath/loops_utils.h.src
// pairwise_sum: blocks of 128, 8 accumulators) so that `np.sum(col) <= peak_limit`
// compares the same number.
__device__ double np_pairwise_sum(const double* a, int n) {
    if (n < 8) {
        double res = 0.0;
        for (int i = 0; i < n; ++i) res = __dadd_rn(res, a[i]);
        return res;
    } else if (n <= 128) {
        double r[8];
        for (int j = 0; j < 8; ++j) r[j] = a[j];
        int i;
        for (i = 8; i < n - (n % 8); i += 8)
            for (int j = 0; j < 8; ++j) r[j] = __dadd_rn(r[j], a[i + j]);
        double res = __dadd_rn(__dadd_rn(__dadd_rn(r[0], r[1]), __dadd_rn(r[2], r[3])),
                               __dadd_rn(__dadd_rn(r[4], r[5]), __dadd_rn(r[6], r[7])));
        for (; i < n; ++i) res = __dadd_rn(res, a[i]);
        return res;
    } else {
        int n2 = n / 2;
        n2 -= n2 % 8;
        return __dadd_rn(np_pairwise_sum(a, n2), np_pairwise_sum(a + n2, n - n2));
    }
}

// all SOC line currents of `col` (length N) within limit + 1e-7; warp-cooperative,
// each lane takes constraint rows j = lane, lane+32, ...
__device__ bool warp_feasible(const SiteDev& S, const double* col, int lane) {
    bool ok = true;
    for (int j = lane; j < S.M; j += 32) {
        double x = 0.0, y = 0.0;
        const double* ac = S.a_cos + (size_t)j * S.N;
        const double* as = S.a_sin + (size_t)j * S.N;
        for (int i = 0; i < S.N; ++i) {
            x = __dadd_rn(x, __dmul_rn(ac[i], col[i]));
            y = __dadd_rn(y, __dmul_rn(as[i], col[i]));
        }
        double cur = __dsqrt_rn(__dadd_rn(__dmul_rn(x, x), __dmul_rn(y, y)));
        if (!(cur <= __dadd_rn(S.limits[j], 1e-7))) ok = false;
    }
    return __all_sync(0xffffffffu, ok);
}

// one warp per instance
__global__ void k_reallocate(SiteDev S, int mode, const double* rates_in, double* rates, int B, int T, int S_max,
                             const int32_t* n_sessions, const int32_t* sess_row, const int32_t* sess_start,
                             const double* sess_ramp, const double* sess_max0, const int32_t* order_in,
                             const double* peak_limit_in) {
    extern __shared__ __align__(16) unsigned char smraw[];
    const int b = blockIdx.x, lane = threadIdx.x, N = S.N;
    double* col = reinterpret_cast<double*>(smraw);       // [N] current first column
    double* trial = col + N;                               // [N]
    double* ub = trial + N;                                // [N]
    double* metric = ub + N;                               // [S_max]
    int* order = reinterpret_cast<int*>(metric + S_max);   // [S_max] -> EVSE index per sorted session
    int* active = order + S_max;                           // [N]
    const int nS = n_sessions[b];
    double* R = rates + (size_t)b * N * T;
    const double* Rin = rates_in + (size_t)b * N * T;
    double peak_limit = 0.0;
    // init_rates / peak limit come from the *continuous* schedule (pp.py:209-211)
    for (int i = lane; i < N; i += 32) { trial[i] = Rin[(size_t)i * T]; col[i] = R[(size_t)i * T]; active[i] = 0; ub[i] = 0.0; }
    __syncwarp();
    if (lane == 0) {
        peak_limit = (mode == 1) ? np_pairwise_sum(trial, N) : peak_limit_in[b];
        for (int s = 0; s < nS; ++s) {
            size_t k = (size_t)b * S_max + s;
            int i = sess_row[k];
            if (sess_start[k] == 0) {  // pp.py:154-164 / 226-236
                active[i] = 1;
                double u = sess_ramp[k];
                if (sess_max0[k] < u) u = sess_max0[k];
                if (S.max_pilot[i] < u) u = S.max_pilot[i];
                ub[i] = u;
            }
            if (mode == 1) metric[s] = -(trial[i] - col[i]);  // pp.py:214-216
            order[s] = (mode == 1) ? s : order_in[k];
        }
        if (mode == 1) {
            // stable insertion sort by metric (Python sorted() is stable)
            for (int a = 1; a < nS; ++a) {
                int oa = order[a];
                double ma = metric[oa];
                int p = a - 1;
                while (p >= 0 && metric[order[p]] > ma) { order[p + 1] = order[p]; --p; }
                order[p + 1] = oa;
            }
        }
        for (int s = 0; s < nS; ++s) order[s] = sess_row[(size_t)b * S_max + order[s]];
    }
    peak_limit = __shfl_sync(0xffffffffu, peak_limit, 0);
    __syncwarp();
    int n_active = 0;
    for (int i = lane; i < N; i += 32) n_active += active[i];
    n_active = __reduce_add_sync(0xffffffffu, n_active);
    // for i in cycle(sorted_indexes)
    int ptr = 0, skipped = 0;
    while (nS > 0 && n_active > 0) {
        const int i = order[ptr];
        ptr = (ptr + 1 == nS) ? 0 : ptr + 1;
        if (!active[i]) {
            if (++skipped >= nS) break;  // an active EVSE that no listed session refers to
            continue;
        }
        skipped = 0;
        const double cur = col[i];
        bool accept = false, deact = false;
        if (cur >= ub[i]) deact = true;
        else {
            int o = S.allow_off[i];
            const double nv = increment_in_set(cur, S.allow_vals + o, S.allow_off[i + 1] - o);
            for (int q = lane; q < N; q += 32) trial[q] = (q == i) ? nv : col[q];
            __syncwarp();
            double sum = 0.0;
            if (lane == 0) sum = np_pairwise_sum(trial, N);
            sum = __shfl_sync(0xffffffffu, sum, 0);
            bool feas = warp_feasible(S, trial, lane);
            // nv == cur cannot make progress (the reference would spin forever here): stop this EVSE
            accept = (sum <= peak_limit) && (nv <= ub[i]) && feas && (nv != cur);
            if (accept) { if (lane == 0) col[i] = nv; } else deact = true;
        }
        if (deact) { if (lane == 0) active[i] = 0; --n_active; }
        __syncwarp();
    }
    for (int i = lane; i < N; i += 32) R[(size_t)i * T] = col[i];
}

__global__ void k_feasible(SiteDev S, const double* rates, int B, int T, int colidx, int32_t* feasible) {
    extern __shared__ __align__(16) unsigned char smraw[];
    double* col = reinterpret_cast<double*>(smraw);
    const int b = blockIdx.x, lane = threadIdx.x;
    for (int i = lane; i < S.N; i += 32) col[i] = rates[((size_t)b * S.N + i) * T + colidx];
    __syncwarp();
    bool ok = warp_feasible(S, col, lane);
    if (lane == 0) feasible[b] = ok ? 1 : 0;
}

// Preprocessing greedy of apply_minimum_charging_rate (acnportal, called at adacharge.py:149-150): sessions are
// offered in the caller's order; a session keeps its minimum rate if the network stays feasible with everything
// admitted before it.  One warp per instance.
__global__ void k_min_rate_admission(SiteDev S, int B, int S_max, const int32_t* n_sessions, const int32_t* sess_row,
                                     const double* try_rate, int32_t* admitted) {
    extern __shared__ __align__(16) unsigned char smraw[];
    double* col = reinterpret_cast<double*>(smraw);
    const int b = blockIdx.x, lane = threadIdx.x;
    for (int i = lane; i < S.N; i += 32) col[i] = 0.0;
    __syncwarp();
    const int nS = n_sessions[b];
    for (int s = 0; s < nS; ++s) {
        const size_t k = (size_t)b * S_max + s;
        const int i = sess_row[k];
        if (lane == 0) col[i] = try_rate[k];
        __syncwarp();
        const bool ok = warp_feasible(S, col, lane);
        if (lane == 0) {
            if (!ok) col[i] = 0.0;
            admitted[k] = ok ? 1 : 0;
        }
        __syncwarp();
    }
}

static int grid_for(size_t total) { return (int)std::min<size_t>((total + 255) / 256, 148 * 16); }

extern "C" int acb_project_continuous(acb_site* site, const double* in, double* out, int B, int T, void* stream) {
    if (!site || !in || !out || B <= 0 || T <= 0) { acb_set_error("acb_project_continuous: bad arguments"); return ACB_E_INVALID; }
    ACB_CUDA(cudaSetDevice(site->device));
    k_project_continuous<<<grid_for((size_t)B * site->d.N * T), 256, 0, (cudaStream_t)stream>>>(site->d, in, out, B, T);
    ACB_CUDA(cudaGetLastError());
    return ACB_OK;
}

extern "C" int acb_project_discrete(acb_site* site, const double* in, double* out, int B, int T, void* stream) {
    if (!site || !in || !out || B <= 0 || T <= 0) { acb_set_error("acb_project_discrete: bad arguments"); return ACB_E_INVALID; }
    if (site->d.nAllow == 0) { acb_set_error("acb_project_discrete: site has no allowable_pilots"); return ACB_E_INVALID; }
    ACB_CUDA(cudaSetDevice(site->device));
    k_project_discrete<<<grid_for((size_t)B * site->d.N * T), 256, 0, (cudaStream_t)stream>>>(site->d, in, out, B, T);
    ACB_CUDA(cudaGetLastError());
    return ACB_OK;
}

extern "C" int acb_reallocate(acb_site* site, int mode, const double* rates_in, double* rates_out, int B, int T,
                              int S_max, const int32_t* n_sessions, const int32_t* sess_row, const int32_t* sess_start,
                              const double* sess_ramp, const double* sess_max0, const int32_t* order,
                              const double* peak_limit, void* stream) {
    if (!site || !rates_in || !rates_out || B <= 0 || T <= 0 || S_max <= 0 || (mode != 0 && mode != 1) ||
        (mode == 0 && (!order || !peak_limit))) {
        acb_set_error("acb_reallocate: bad arguments");
        return ACB_E_INVALID;
    }
    if (site->d.nAllow == 0) { acb_set_error("acb_reallocate: site has no allowable_pilots"); return ACB_E_INVALID; }
    ACB_CUDA(cudaSetDevice(site->device));
    cudaStream_t st = (cudaStream_t)stream;
    const size_t total = (size_t)B * site->d.N * T;
    if (mode == 1) {
        k_project_discrete<<<grid_for(total), 256, 0, st>>>(site->d, rates_in, rates_out, B, T);
        ACB_CUDA(cudaGetLastError());
    } else if (rates_in != rates_out) {
        ACB_CUDA(cudaMemcpyAsync(rates_out, rates_in, total * sizeof(double), cudaMemcpyDeviceToDevice, st));
    }
    size_t smem = (size_t)(3 * site->d.N + S_max) * sizeof(double) + (size_t)(S_max + site->d.N) * sizeof(int);
    if (smem > 48 * 1024) ACB_CUDA(cudaFuncSetAttribute(k_reallocate, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    k_reallocate<<<B, 32, smem, st>>>(site->d, mode, rates_in, rates_out, B, T, S_max, n_sessions, sess_row, sess_start,
                                      sess_ramp, sess_max0, order, peak_limit);
    ACB_CUDA(cudaGetLastError());
    return ACB_OK;
}

extern "C" int acb_constraints_feasible(acb_site* site, const double* rates, int B, int T, int col, int32_t* feasible, void* stream) {
    if (!site || !rates || !feasible || B <= 0 || col < 0 || col >= T) { acb_set_error("acb_constraints_feasible: bad arguments"); return ACB_E_INVALID; }
    ACB_CUDA(cudaSetDevice(site->device));
    k_feasible<<<B, 32, site->d.N * sizeof(double), (cudaStream_t)stream>>>(site->d, rates, B, T, col, feasible);
    ACB_CUDA(cudaGetLastError());
    return ACB_OK;
}

extern "C" int acb_min_rate_admission(acb_site* site, int B, int S_max, const int32_t* n_sessions, const int32_t* sess_row,
                                      const double* try_rate, int32_t* admitted, void* stream) {
    if (!site || !n_sessions || !sess_row || !try_rate || !admitted || B <= 0 || S_max <= 0) {
        acb_set_error("acb_min_rate_admission: bad arguments");
        return ACB_E_INVALID;
    }
    ACB_CUDA(cudaSetDevice(site->device));
    k_min_rate_admission<<<B, 32, site->d.N * sizeof(double), (cudaStream_t)stream>>>(site->d, B, S_max, n_sessions, sess_row, try_rate, admitted);
    ACB_CUDA(cudaGetLastError());
    return ACB_OK;
}
