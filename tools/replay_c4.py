"""BASELINE config 4 at scale: closed-loop warm-started replay of many Caltech-shaped sites.
  python tools/replay_c4.py [n_sites] [t0] [t1] [days] [Tp] [host|device]   (one GPU; default: simulator step on the device)
  torchrun --nproc-per-node N tools/replay_c4.py ...                   (sites sharded over ranks)
Prints one JSON line: control steps per second for the whole fleet, site-steps per second,
the host/device split and the iteration statistics."""
import json
import os
import sys
import time

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
import numpy as np
import torch

import adacharge_b200 as ab
from adacharge_b200.generators import caltech_acn_infrastructure
from adacharge_b200.replay_fast import DeviceFleetReplay, FleetReplay
from adacharge_b200 import sharding

n_sites = int(sys.argv[1]) if len(sys.argv) > 1 else 1024
t0 = int(sys.argv[2]) if len(sys.argv) > 2 else 0
t1 = int(sys.argv[3]) if len(sys.argv) > 3 else 288
days = int(sys.argv[4]) if len(sys.argv) > 4 else 1
Tp = int(sys.argv[5]) if len(sys.argv) > 5 else 160
on_device = (sys.argv[6] if len(sys.argv) > 6 else "device") == "device"
rank, world = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1))
local = int(os.environ.get("LOCAL_RANK", 0))
torch.cuda.set_device(local)
if world > 1:
    if os.environ.get("NCCL_DEBUG", "VERSION").upper() == "VERSION":
        os.environ["NCCL_DEBUG"] = "WARN"
    os.environ.setdefault("NCCL_DEBUG_FILE", "/dev/stderr")
    torch.distributed.init_process_group("nccl", device_id=torch.device("cuda", local))
_r = sharding.shard_range(n_sites, rank, world)
lo, hi = _r.start, _r.stop
obj = [ab.ObjectiveComponent(ab.tou_energy_cost), ab.ObjectiveComponent(ab.total_energy, 0.3), ab.ObjectiveComponent(ab.demand_charge, 1 / 30)]
rp = (DeviceFleetReplay if on_device else FleetReplay)(caltech_acn_infrastructure(), obj, n_sites=hi - lo, steps_per_day=288, days=days, seed0=1000, Tp=Tp, site_offset=lo,
                 solver_options=json.loads(os.environ.get("ACB_REPLAY_OPTS", "{}")))
rp.run(t0, min(t0 + 3, t1))  # warm-up steps (library load, allocator), not timed
if world > 1:
    torch.distributed.barrier()
torch.cuda.synchronize()
w0 = time.perf_counter()
n_before = len(rp.stats.device_ms)
s_before = rp.summary()
stats = rp.run(t0 + 3, t1)
torch.cuda.synchronize()
wall = time.perf_counter() - w0
s = rp.summary()
if on_device:
    timed = dict(steps=t1 - t0 - 3, device_ms=wall * 1e3, host_ms=0.0, site_steps=s["site_steps"] - s_before["site_steps"])
    s = dict(s, unsolved=s["unsolved"] - s_before["unsolved"], iters_max=float("nan"))
else:
    timed = dict(steps=len(stats.device_ms) - n_before, device_ms=sum(stats.device_ms[n_before:]), host_ms=sum(stats.host_ms[n_before:]),
                 site_steps=sum(stats.active_sites[n_before:]))
if world > 1:
    t = torch.tensor([wall, timed["device_ms"], timed["host_ms"]], device="cuda", dtype=torch.float64)
    torch.distributed.all_reduce(t, op=torch.distributed.ReduceOp.MAX)
    c = torch.tensor([timed["site_steps"], s["unsolved"], float(np.sum(stats.delivered_frac)), s["iters_max"]], device="cuda", dtype=torch.float64)
    mx = c[3:].clone()
    torch.distributed.all_reduce(c, op=torch.distributed.ReduceOp.SUM)
    torch.distributed.all_reduce(mx, op=torch.distributed.ReduceOp.MAX)
    wall, dms, hms = t.tolist()
    site_steps, unsolved, full = c[:3].tolist()
    it_max = mx.item()
else:
    dms, hms, site_steps, unsolved, full, it_max = timed["device_ms"], timed["host_ms"], timed["site_steps"], s["unsolved"], float(np.sum(stats.delivered_frac)), s["iters_max"]
if rank == 0:
    print(json.dumps(dict(workload="C4 replay", simulator="device" if on_device else "host", n_sites=n_sites, n_gpus=world, steps=timed["steps"], Tp=Tp, days=days,
                          control_steps_per_s=round(timed["steps"] / wall, 2), site_steps_per_s=round(site_steps / wall, 1),
                          wall_s=round(wall, 2), device_ms_per_step=round(dms / max(timed["steps"], 1), 2),
                          host_ms_per_step=round(hms / max(timed["steps"], 1), 2), iters_mean_rank0=round(s["iters_mean"], 1),
                          iters_max=it_max, unsolved=unsolved, mean_delivered_fraction=round(full / n_sites, 4))))
if world > 1:
    torch.distributed.destroy_process_group()
