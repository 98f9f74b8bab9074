// Closed-loop replay on the device: the simulator side of one control step for a fleet of sites.
//
// The reference is driven by acnportal's Simulator, which at every period hands the algorithm the active sessions
// (adacharge/adacharge.py:18-39 get_active_sessions: plugged in, energy still owed) and applies the first-period pilots
// it returns (call pattern adacharge/adacharge.py:135-193; loop shape exercised by tests/test_integration.py:115-118).
// For a fleet replay that loop is two kernels around the solve, with the EV table, the energy delivered so far, the
// previous peak and the per-EV warm-start multipliers resident on the device:
//   acb_fleet_sessions   step t -> the raw [sites][S_max] session tables acb_pack_sessions consumes (+ the EV behind each
//                        slot and its warm-start multiplier)
//   acb_fleet_apply      first-period pilots -> energy delivered, previous peak, per-EV multipliers, step statistics
// The EV table is sorted by (day, site, station), so a site's active sessions come out ordered by EVSE row.
#include "acb_common.cuh"

__global__ void acb_fleet_sessions_kernel(acb_fleet F, int t, acb_sessions X, int32_t* sess_ev, float* warm_mu) {
    const int b = blockIdx.x, lane = threadIdx.x;  // one warp per site
    const int day = min(t / F.steps_per_day, F.days - 1);
    const int lo = F.day_site_off[(size_t)day * F.n_sites + b], hi = F.day_site_off[(size_t)day * F.n_sites + b + 1];
    const size_t base = (size_t)b * X.S_max;
    int32_t* station = const_cast<int32_t*>(X.station) + base;
    int32_t* arr = const_cast<int32_t*>(X.arrival_offset) + base;
    int32_t* rem_t = const_cast<int32_t*>(X.remaining_time) + base;
    double* dem = const_cast<double*>(X.remaining_demand) + base;
    double* mn = const_cast<double*>(X.min_rate) + base;
    double* mx = const_cast<double*>(X.max_rate) + base;
    const bool had = F.had[b] != 0;
    int n = 0;
    for (int e0 = lo; e0 < hi; e0 += 32) {
        const int e = e0 + lane;
        bool act = false;
        double owed = 0.0;
        if (e < hi) {
            owed = F.ev_req[e] - F.ev_dlv[e];
            act = F.ev_arr[e] <= t && t < F.ev_dep[e] && owed > 1e-6;  // plugged in and energy still owed
        }
        const unsigned m = __ballot_sync(0xffffffffu, act);
        const int j = n + __popc(m & ((1u << lane) - 1u));
        if (act && j < X.S_max) {
            station[j] = F.ev_station[e];
            arr[j] = 0;
            rem_t[j] = F.ev_dep[e] - t;
            dem[j] = owed;
            mn[j] = 0.0;
            mx[j] = F.ev_max[e];
            sess_ev[base + j] = e;
            warm_mu[base + j] = had ? F.ev_mu[e] : 0.f;
        }
        n += __popc(m);
    }
    n = min(n, X.S_max);
    for (int j = n + lane; j < X.S_max; j += 32) {
        station[j] = -1; arr[j] = 0; rem_t[j] = 0; dem[j] = 0.0; mn[j] = 0.0; mx[j] = 0.0;
        sess_ev[base + j] = -1;
        warm_mu[base + j] = 0.f;
    }
}

__global__ void acb_fleet_apply_kernel(SiteDev S, acb_fleet F, int t, double period, const double* pilots, int Tp, const int32_t* n_sessions,
                                       const int32_t* sess_ev, int S_max, const float* out_mu, const int32_t* status, const int32_t* iters,
                                       double* first_out, double* stats) {
    const int b = blockIdx.x, lane = threadIdx.x;  // one warp per site
    const int N = S.N;
    const int day = min(t / F.steps_per_day, F.days - 1);
    const int lo = F.day_site_off[(size_t)day * F.n_sites + b], hi = F.day_site_off[(size_t)day * F.n_sites + b + 1];
    const double* p0 = pilots + (size_t)b * N * Tp;  // pilots[b][i][0] = p0[i * Tp]
    // the simulator side: the first-period pilots charge every EV that is plugged in (also those that owe nothing)
    for (int e = lo + lane; e < hi; e += 32) {
        if (F.ev_arr[e] <= t && t < F.ev_dep[e]) {
            const int i = F.ev_station[e];
            const double w = S.volt[i] * period / 1e3 / 60;  // kWh per A*period
            const double owed = F.ev_req[e] - F.ev_dlv[e];
            double de = p0[(size_t)i * Tp] * w;
            de = (owed < de) ? owed : de;
            F.ev_dlv[e] += (de > 0.0) ? de : 0.0;
        }
    }
    const int n = n_sessions[b];
    for (int j = lane; j < n; j += 32) {
        const int e = sess_ev[(size_t)b * S_max + j];
        if (e >= 0) F.ev_mu[e] = out_mu[(size_t)b * S_max + j];
    }
    if (first_out)
        for (int i = lane; i < N; i += 32) first_out[(size_t)b * N + i] = p0[(size_t)i * Tp];
    if (lane == 0) {
        // aggregate first-period current, summed in numpy's pairwise order (the host replay takes first.sum(axis=1))
        double tot;
        if (N < 8) {
            tot = 0.0;
            for (int i = 0; i < N; ++i) tot = __dadd_rn(tot, p0[(size_t)i * Tp]);
        } else {
            double r[8];
            for (int j = 0; j < 8; ++j) r[j] = p0[(size_t)j * Tp];
            int i;
            for (i = 8; i < N - (N % 8); i += 8)
                for (int j = 0; j < 8; ++j) r[j] = __dadd_rn(r[j], p0[(size_t)(i + j) * Tp]);
            tot = __dadd_rn(__dadd_rn(__dadd_rn(r[0], r[1]), __dadd_rn(r[2], r[3])), __dadd_rn(__dadd_rn(r[4], r[5]), __dadd_rn(r[6], r[7])));
            for (; i < N; ++i) tot = __dadd_rn(tot, p0[(size_t)i * Tp]);
        }
        if (tot > F.prev_peak[b]) F.prev_peak[b] = tot;
        F.had[b] = n > 0 ? 1 : 0;
        if (stats && n > 0) {
            atomicAdd(stats + 0, 1.0);                                  // site-steps solved
            atomicAdd(stats + 1, (double)iters[b]);                     // iterations
            if (status[b] != ACB_SOLVED) atomicAdd(stats + 2, 1.0);     // not certified
        }
    }
}

extern "C" int acb_fleet_sessions(acb_site* site, const acb_fleet* fleet, int t, const acb_sessions* sessions, int32_t* sess_ev, float* warm_mu,
                                  void* stream) {
    if (!site || !fleet || !sessions || !sess_ev || !warm_mu || sessions->B != fleet->n_sites || t < 0) {
        acb_set_error("acb_fleet_sessions: bad arguments (one instance per site)");
        return ACB_E_INVALID;
    }
    ACB_CUDA(cudaSetDevice(site->device));
    acb_fleet_sessions_kernel<<<fleet->n_sites, 32, 0, (cudaStream_t)stream>>>(*fleet, t, *sessions, sess_ev, warm_mu);
    ACB_CUDA(cudaGetLastError());
    return ACB_OK;
}

extern "C" int acb_fleet_apply(acb_site* site, const acb_fleet* fleet, int t, double period, const acb_batch* batch, const int32_t* sess_ev,
                               double* first_pilots, double* stats, void* stream) {
    if (!site || !fleet || !batch || !batch->pilots || !batch->out_mu || !sess_ev || batch->B != fleet->n_sites) {
        acb_set_error("acb_fleet_apply: bad arguments (the batch needs pilots and out_mu)");
        return ACB_E_INVALID;
    }
    ACB_CUDA(cudaSetDevice(site->device));
    acb_fleet_apply_kernel<<<fleet->n_sites, 32, 0, (cudaStream_t)stream>>>(site->d, *fleet, t, period, batch->pilots, batch->Tp, batch->n_sessions,
                                                                             sess_ev, batch->S_max, batch->out_mu, batch->status, batch->iters,
                                                                             first_pilots, stats);
    ACB_CUDA(cudaGetLastError());
    return ACB_OK;
}
