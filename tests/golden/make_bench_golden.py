"""Oracle optima of a sample of the benchmark workload (BASELINE.json configs[2], bench.py's C3 batch).

    python tests/golden/make_bench_golden.py [--stride 16] [--count 256] [--procs 6]

Instance b of the batch is ``config_c2(seed=b, price_noise=0.2)`` with the objective of bench.py
(tou_energy_cost + 0.3 total_energy + demand_charge / 30).  Every ``stride``-th instance is solved by the CPU oracle
(oracle/mpc.py, float64 interior point) and its optimal objective is stored, so that bench.py can gate the timed run
on a sample spread over the whole batch (objective within 1e-4 |f*|, violation, energy, bounds) without paying for
oracle solves at bench time.  Output: tests/golden/bench_c3_golden.json.
"""
import argparse
import json
import multiprocessing as mp
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)

BENCH_OBJECTIVE = [("tou_energy_cost", 1.0, {}), ("total_energy", 0.3, {}), ("demand_charge", 1.0 / 30.0, {})]


def one(seed):
    for var in ("OMP_NUM_THREADS", "OPENBLAS_NUM_THREADS", "MKL_NUM_THREADS"):
        os.environ[var] = "1"
    from adacharge_b200.interface import TestingInterface
    from adacharge_b200.generators import config_c2
    from oracle import mpc

    iface = TestingInterface(config_c2(seed, price_noise=0.2))
    S, I = iface.active_sessions(), iface.infrastructure_info()
    t = time.perf_counter()
    R, info = mpc.solve_mpc(BENCH_OBJECTIVE, S, I, iface, "SOC", False, None, iface.get_prev_peak(), return_info=True)
    f = float(mpc.evaluate_objective(R, BENCH_OBJECTIVE, I, iface, S, iface.get_prev_peak()))
    v = mpc.violations(R, S, I, iface)
    return dict(seed=seed, objective=f, oracle_iters=int(info["iters"]), oracle_gap=float(info["gap"]),
                oracle_violation=float(max(v["infrastructure_rel"], 0.0)), n_sessions=len(S), seconds=time.perf_counter() - t)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--stride", type=int, default=16)
    ap.add_argument("--count", type=int, default=256)
    ap.add_argument("--procs", type=int, default=max(1, (os.cpu_count() or 2) - 2))
    ap.add_argument("--out", default=os.path.join(ROOT, "tests", "golden", "bench_c3_golden.json"))
    a = ap.parse_args()
    seeds = [i * a.stride for i in range(a.count)]
    t = time.perf_counter()
    with mp.get_context("spawn").Pool(a.procs) as pool:
        rows = pool.map(one, seeds, chunksize=1)
    out = dict(workload="config_c2(seed, price_noise=0.2), SOC, T=288", objective=[[n, c, k] for n, c, k in BENCH_OBJECTIVE],
               stride=a.stride, instances=rows, wall_seconds=time.perf_counter() - t, procs=a.procs)
    with open(a.out, "w") as f:
        json.dump(out, f, indent=0)
    print(f"{len(rows)} instances in {out['wall_seconds']:.0f} s -> {a.out}")


if __name__ == "__main__":
    main()
