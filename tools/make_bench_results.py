"""DEV (CPU only): assemble profiles/r02_bench_results.md from the JSON lines that the gpurun calls left in gpurun_out/."""
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
G = os.path.join(ROOT, "gpurun_out")


def load(name):
    p = os.path.join(G, name)
    if not os.path.exists(p):
        return None
    try:
        txt = [l for l in open(p).read().splitlines() if l.startswith("{")]
        return json.loads(txt[-1]) if txt else None
    except Exception:
        return None


def short(l, keys):
    return {k: (round(v, 4) if isinstance(v, float) else v) for k, v in l.items() if k in keys}


out = ["# Round-2 benchmark results (B200, gpurun boxes, never under a profiler)", "",
       "Every line below is the JSON that `bench.py` (or `tools/replay_c4.py`) printed; commands are given with each block.",
       "`value` = solves/s with the raw session tables resident in HBM (device packer + solve + fused pilot projection inside the timed region),",
       "`e2e` = host session tables -> host float64 pilots through `BatchedAdaptiveCharging.schedule_async`.", ""]
KEYS = ("value", "unit", "n_gpus", "steps", "ms_per_step", "scaling", "solved", "instances", "iters_mean", "iters_max", "gpu_launches", "site_groups_per_gpu",
        "host_enqueue_ms_per_step_max_rank")
blocks = [
    ("C3 (headline), 1 GPU: `python bench.py`", "r02x_bench.json"),
    ("C3, 2 GPUs, weak: `torchrun --nproc-per-node 2 bench.py --gpus 2`", "r02u_n2_c3.json"),
    ("C3, 2 GPUs, strong (4096 instances in total): `... bench.py --gpus 2 --scaling strong`", "r02u_n2_c3s.json"),
    ("C3, 8 GPUs, weak: `torchrun --nproc-per-node 8 bench.py --gpus 8`", "r02_n8_c3.json"),
    ("C3, 8 GPUs, strong: `... bench.py --gpus 8 --scaling strong`", "r02_n8_c3s.json"),
    ("C1 (one instance per step): `python bench.py --config c1`", "r02_c1_bench.json"),
    ("C2 (one instance per step): `python bench.py --config c2`", "r02_c2_bench.json"),
    ("C4 (closed-loop replay on the device, 1024 sites, 40 control steps from 8 am), 1 GPU: `python bench.py --config c4 --steps 40`", "r02_c4_bench.json"),
    ("C4, 2 GPUs (one lockstep group per GPU at the time of this run): `... bench.py --gpus 2 --config c4 --steps 40`", "r02u_n2_c4.json"),
    ("C4, 8 GPUs (2 groups of 64 sites per GPU): `... bench.py --gpus 8 --config c4 --steps 40 --groups 2`", "r02_n8_c4.json"),
    ("C4, 8 GPUs (8 groups of 16 sites per GPU, the default): `... bench.py --gpus 8 --config c4 --steps 40`", "r02E_n8_c4.json"),
    ("C5 (1000 EVSEs, general path, 128 instances per GPU), 1 GPU: `python bench.py --config c5 --steps 4`", "r02s_c5.json"),
    ("C5, 2 GPUs: `... bench.py --gpus 2 --config c5 --steps 4`", "r02u_n2_c5.json"),
    ("C5, 8 GPUs: `... bench.py --gpus 8 --config c5 --steps 4`", "r02_n8_c5.json"),
    ("Reference arm (CPU oracle on every host core; 24-core box): `bench.py --impl reference --steps 3 --warmup 1`", "r02u_n2_ref.json"),
]
for title, f in blocks:
    l = load(f)
    if l is None:
        continue
    out += [f"## {title}", "", "```json", json.dumps(short(l, KEYS))]
    for k in ("e2e", "roofline", "parity_sample", "latency", "cpu_baseline", "clocks"):
        if k in l and l[k]:
            v = l[k]
            if k == "parity_sample":
                v = {a: b for a, b in v.items() if a not in ("source",)}
            if k == "roofline":
                v = {a: b for a, b in v.items() if a not in ("note", "peak_source")}
            out.append(json.dumps({k: v}))
    out += ["```", ""]
out += ["## Closed-loop replay, full BASELINE config 4 (`tools/replay_c4.py n_sites t0 t1 days Tp device groups`)", "", "```json"]
for f in ("r02_c4_1day.json", "r02_c4_30d.json", "r02_c4_30d_g1.json", "r02D_30d_g32.json", "r02_n8_c4_30d.json", "r02E_n8_c4_30d.json"):
    l = load(f)
    if l:
        out.append(json.dumps(l))
out += ["```", "",
        "30 days x 1024 sites x 288 steps = 5.07 M warm-started MPC solves: **11.1 s on one B200** with the simulator step on the device and the",
        "sites in 8 independent groups per GPU (20.0 s in one lockstep group, 22.7 s in 32 groups; round 1: 39.0 s with the host-side simulator",
        "step).  One GPU is throughput-bound at 8 groups (the step time equals the summed block time / 148 SMs).  On 8 GPUs the same fleet takes 4.6 s",
        "(6.3 s with 2 groups per GPU; round 1: 21.5 s): with 128 sites per GPU a control step is bound by the latency of its slowest site",
        "(up to ~700 iterations x 10 us), not by throughput, so a fixed 1024-site fleet does not scale further.", "",
        "C4 bench, groups per GPU on one GPU (`bench.py --config c4 --steps 40 --groups G`): 1: 77.7 k, 4: 104.3 k, 8: 110.5-115.2 k, 16: 71.0 k, 32: 50.1 k,",
        "64: 32.8 k site-steps/s.", ""]
open(os.path.join(ROOT, "profiles", "r02_bench_results.md"), "w").write("\n".join(out))
print("\n".join(out)[:1500])
