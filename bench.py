#!/usr/bin/env python
"""MPC solves/sec on the batched CaltechACN three-phase workload (BASELINE.json configs[2]).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--batch B] [--impl reference]

A step = one pass of the hot path (build -> solve, AdaptiveChargingOptimization.solve for
every instance of the batch) over one batch of B synthetic C2-shaped instances per GPU
(weak scaling; instances are independent, no collective on the solve path, one final
gather).  `value` is timed with the packed inputs already resident in HBM; `e2e` times
pinned host staging -> H2D -> solve -> D2H of the schedules.  The `--impl reference` arm
times the CPU oracle (the float64 restatement of the reference's cvxpy/ECOS path; the
reference itself cannot run here: cvxpy/ECOS/acnportal are not installed) on the host
cores, one instance per core per step.
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

import numpy as np  # noqa: E402

METRIC = "mpc_solves_per_sec_1e-4_rel_gap"
UNIT = "solves/s"
BENCH_OBJECTIVE = [("tou_energy_cost", 1.0, {}), ("total_energy", 0.3, {}), ("demand_charge", 1.0 / 30.0, {})]
WORKLOAD = ("C3: batch of independent CaltechACN three-phase 54-EVSE MPC instances, SOC constraints, T=288, "
            "tou_energy_cost + 0.3*total_energy + (1/30)*demand_charge, randomised sessions/prices")


def make_config(args, n_gpus):
    return {
        "workload": WORKLOAD, "instances_per_gpu": args.batch, "global_instances": args.batch * n_gpus, "N": 54, "T": 288, "M": 8,
        "parallelism": f"independent instances sharded over {n_gpus} GPU(s), no solve-path collective",
        "tolerances": {"eps_rel": 1e-4, "eps_abs": 1e-5, "violation": 1e-5},
        "l2": "256 MiB buffer written between timed steps (outside the event-timed region); per-step output 54x288xB fp32 exceeds L2 at B>=2048",
    }


def build_instances(batch, seed0):
    import adacharge_b200 as ab
    from adacharge_b200.generators import config_c2, caltech_acn_infrastructure

    infra = caltech_acn_infrastructure()
    obj = [ab.ObjectiveComponent(getattr(ab, n), c, k) for n, c, k in BENCH_OBJECTIVE]
    insts, site, ifaces = [], None, []
    for i in range(batch):
        d = config_c2(seed0 + i, infra=infra, price_noise=0.2)
        iface = ab.TestingInterface(d)
        S, I = iface.active_sessions(), iface.infrastructure_info()
        aco = ab.AdaptiveChargingOptimization(obj, iface)
        inst = aco.build_instance(S, I, None, iface.get_prev_peak())
        insts.append(inst)
        if site is None:
            site = aco._site_for(I, inst)
        if i < 64:
            ifaces.append(iface)
    return site, insts, ifaces


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled during the timed region."""

    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.index, self.rows, self.proc = index, [], None

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.index), f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-lms", "200"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.thread = threading.Thread(target=self._read, daemon=True)
            self.thread.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([c.strip() for c in line.split(",")])

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)
        except Exception:
            self.proc.kill()
        sm = [float(r[0]) for r in self.rows if r and r[0].replace(".", "").isdigit()]
        mx = [float(r[1]) for r in self.rows if len(r) > 1 and r[1].replace(".", "").isdigit()]
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        reasons = sorted({n for r in self.rows if len(r) >= 7 for n, v in zip(names, r[3:7]) if v.lower().startswith("active")})
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None, "reasons": reasons, "samples": len(sm)}


def _oracle_one(seed):
    os.environ.setdefault("OMP_NUM_THREADS", "1")
    from adacharge_b200.interface import TestingInterface
    from adacharge_b200.generators import config_c2
    from oracle import mpc

    iface = TestingInterface(config_c2(seed, price_noise=0.2))
    S, I = iface.active_sessions(), iface.infrastructure_info()
    t = time.perf_counter()
    R = mpc.solve_mpc(BENCH_OBJECTIVE, S, I, iface, "SOC", False, None, iface.get_prev_peak())
    return time.perf_counter() - t, float(mpc.evaluate_objective(R, BENCH_OBJECTIVE, I, iface, S, iface.get_prev_peak()))


def cpu_oracle_throughput(n_instances, cores, seed0=0):
    """Solves `n_instances` instances of the workload with the CPU oracle, one per worker."""
    import multiprocessing as mp

    # one single-threaded solver per core: set before the workers import numpy/scipy
    for var in ("OMP_NUM_THREADS", "OPENBLAS_NUM_THREADS", "MKL_NUM_THREADS"):
        os.environ[var] = "1"
    ctx = mp.get_context("spawn")
    t = time.perf_counter()
    with ctx.Pool(cores) as pool:
        res = pool.map(_oracle_one, [seed0 + i for i in range(n_instances)])
    wall = time.perf_counter() - t
    return n_instances / wall, wall, res


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    cores = max(1, min(os.cpu_count() or 1, 16))
    # bounded sample: one instance per worker, fewer workers per step when many steps are asked
    # for, so that the whole run stays within a few minutes (an oracle solve takes 10-40 s)
    per_step = cores if args.steps <= 6 else int(min(cores, max(2, (cores * 6) // max(args.steps, 1))))
    if args.warmup > 0:  # one small warm-up pass is enough to page the interpreter in
        cpu_oracle_throughput(2, 2, seed0=10_000)
    t = time.perf_counter()
    n = 0
    for k in range(args.steps):
        cpu_oracle_throughput(per_step, per_step, seed0=k * per_step)
        n += per_step
    wall = time.perf_counter() - t
    value = n / wall
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": wall / args.steps * 1e3, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64",
        "data": "synthetic", "config": make_config(args, args.gpus),
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": per_step, "kind": "port",
                         "sample": f"{per_step} instances of the workload per step (one per core), oracle/mpc.py interior-point restatement; "
                                   "the reference's cvxpy/ECOS path cannot run here (cvxpy, ecos, acnportal not installed)"},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--batch", type=int, default=4096, help="instances per GPU per step")
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--no-cpu-baseline", action="store_true")
    args = ap.parse_args()
    if args.impl == "reference":
        return run_reference(args)

    import torch
    import torch.distributed as dist

    from adacharge_b200 import _cabi, engine, sharding

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    assert torch.cuda.is_available(), "bench.py needs a CUDA device"
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        # rank 0 prints exactly one JSON line on stdout.  NCCL writes its version banner to stdout when NCCL_DEBUG is
        # VERSION (the GPU boxes export that) and to NCCL_DEBUG_FILE at WARN/INFO: raise VERSION to WARN, point the
        # log at stderr, and keep fd 1 on stderr while the communicator comes up.
        if os.environ.get("NCCL_DEBUG", "VERSION").upper() == "VERSION":
            os.environ["NCCL_DEBUG"] = "WARN"
        os.environ.setdefault("NCCL_DEBUG_FILE", "/dev/stderr")
        sys.stdout.flush()
        saved_fd = os.dup(1)
        os.dup2(2, 1)
        try:
            dist.init_process_group("nccl", device_id=dev)
            dist.barrier()
            torch.cuda.synchronize()
        finally:
            sys.stdout.flush()
            os.dup2(saved_fd, 1)
            os.close(saved_fd)
    _cabi.lib()

    site, insts, ifaces = build_instances(args.batch, seed0=rank * args.batch)
    opt = _cabi.default_options()
    pb = engine.PackedBatch(site, insts).upload()
    flush = torch.empty(256 * 1024 * 1024 // 4, dtype=torch.float32, device=dev)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def timed(fn, steps):
        tot = 0.0
        for _ in range(steps):
            flush.fill_(1.0)
            torch.cuda.synchronize()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            fn()
            e1.record()
            torch.cuda.synchronize()
            tot += e0.elapsed_time(e1)
        return tot  # ms

    def step_resident():
        pb.solve(opt)

    # end to end through the public host-to-host call: pinned host inputs -> device -> solve -> pinned host results,
    # in 4 chunks on their own streams so the copies overlap the solves (engine.HostPipeline)
    pipe = engine.HostPipeline(site, insts, chunks=4)

    def step_e2e():
        pipe.run(opt)

    for _ in range(max(args.warmup, 3)):
        step_resident()
    barrier()
    clocks = ClockSampler(local)
    if rank == 0:
        clocks.start()
    ms = timed(step_resident, args.steps)
    barrier()
    clk = clocks.stop() if rank == 0 else None
    ms = sharding.max_over_ranks(ms, dev)
    step_e2e()
    barrier()
    ms_e2e = sharding.max_over_ranks(timed(step_e2e, args.steps), dev)
    barrier()

    status = pb.status.cpu().numpy()
    iters = pb.iters.cpu().numpy().astype(np.float64)
    summary = torch.tensor([float((status == 0).sum()), float(len(status)), iters.sum(), iters.max()], dtype=torch.float64, device=dev)
    gathered = sharding.gather_summaries(summary)  # the final gather
    tot = torch.stack(gathered).cpu().numpy()
    n_total = tot[:, 1].sum()
    value = n_total * args.steps / (ms / 1e3)
    e2e = n_total * args.steps / (ms_e2e / 1e3)
    # roofline of the solve kernel: algorithmic bytes per SURVEY.md §8(d) (fp32 state streamed once per iteration)
    N, T, M = 54, 288, 8
    b_iter = 4 * (4 * N * T + 2 * (2 * M + 2) * T)
    alg_bytes_rank = float(((iters + 1) * b_iter).sum())
    launch_ms = ms / args.steps
    peaks = {}
    try:
        peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
    except Exception:
        pass
    peak = float(peaks.get("hbm_gbs", 6650.0))
    achieved = alg_bytes_rank / (launch_ms / 1e3) / 1e9
    traffic = None
    try:
        # measured DRAM bytes per instance (ncu, profiles/) x instances per launch on this rank
        traffic = json.load(open(os.path.join(ROOT, "profiles", "solve_kernel_dram.json")))["dram_bytes_per_instance"] * args.batch
    except Exception:
        pass
    if rank == 0:
        line = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": max(args.warmup, 3),
            "ms_per_step": launch_ms, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32",
            "data": "synthetic", "config": make_config(args, world),
            "solved": int(tot[:, 0].sum()), "instances": int(n_total), "iters_mean": float(tot[:, 2].sum() / n_total), "iters_max": float(tot[:, 3].max()),
            "roofline": {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak, "traffic": traffic,
                         "peak_source": "MEASURED_PEAKS.json hbm_gbs" if peaks else "fallback 6650 (B200_PROFILING.md)",
                         "kernel": "acb_solve_kernel", "algorithmic_bytes_per_iteration": b_iter,
                         "note": "effective bandwidth: state is on-chip resident, DRAM sees load/store only (SURVEY.md 8(d)); per-GPU figure"},
            "e2e": {"value": e2e, "unit": UNIT, "h2d_bytes_per_step": int(pipe.h2d_bytes) * world,
                    "d2h_bytes_per_step": int(pipe.d2h_bytes) * world, "ms_per_step": ms_e2e / args.steps},
            "gpu_launches": args.steps, "clocks": clk,
        }
        if world == 1 and not args.no_cpu_baseline:
            cores = max(1, min(os.cpu_count() or 1, 16))
            v, wall, res = cpu_oracle_throughput(cores, cores)
            line["cpu_baseline"] = {"value": v, "unit": UNIT, "cores": cores, "kind": "port",
                                    "sample": f"first {cores} instances of the batch, one per core, oracle/mpc.py (float64 interior point), wall {wall:.1f} s"}
            # parity of the timed run on that sample: objective within 1e-4 of the oracle
            from oracle import mpc

            rates = pb.rates[:cores].cpu().numpy().astype(np.float64)
            rel = []
            for i in range(min(cores, len(ifaces))):
                S, I = ifaces[i].active_sessions(), ifaces[i].infrastructure_info()
                f = mpc.evaluate_objective(rates[i][:, : insts[i].T], BENCH_OBJECTIVE, I, ifaces[i], S, ifaces[i].get_prev_peak())
                rel.append(abs(f - res[i][1]) / max(abs(res[i][1]), 1e-12))
            line["parity_sample"] = {"max_rel_objective_error": float(max(rel)), "instances": len(rel)}
        print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
