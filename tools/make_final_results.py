"""DEV (CPU only): assemble profiles/r02_final_bench_results.md from the JSON lines that the last gpurun calls of the round left in
gpurun_out/ (tools/gpu_round_evidence.sh for one GPU; the torchrun commands quoted in each block for 2 and 8)."""
import json
import os

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
G = os.path.join(ROOT, "gpurun_out")


def load(name):
    p = os.path.join(G, name)
    if not os.path.exists(p):
        return None
    txt = [l for l in open(p).read().splitlines() if l.startswith("{")]
    return json.loads(txt[-1]) if txt else None


def short(l, keys):
    return {k: (round(v, 4) if isinstance(v, float) else v) for k, v in l.items() if k in keys}


KEYS = ("value", "unit", "n_gpus", "steps", "ms_per_step", "scaling", "solved", "instances", "iters_mean", "iters_max", "gpu_launches", "site_groups_per_gpu",
        "host_enqueue_ms_per_step_max_rank")
TR = "python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port P"
out = ["# Round-2 final benchmark results (B200, gpurun boxes, never under a profiler)", "",
       "Final build of the round (HEAD at the time: \"General path row pass in 2-warp blocks\"; the C3 kernel is the one of commit \"Rate polish: "
       "extrapolation only with strong contraction ...\").  `profiles/r02_bench_results.md` holds the same table for the mid-round build "
       "(on-chip kernel 11 % slower, general path 2.7x slower) plus the 8-GPU 30-day replay and the groups-per-GPU sweep, which were not re-run.",
       "Every line is the JSON that `bench.py` printed.  `value` = solves/s with the raw session tables resident in HBM (device packer + solve + fused "
       "pilot projection inside the timed region), `e2e` = host session tables -> host float64 pilots through "
       "`BatchedAdaptiveCharging.schedule_async`.  Multi-GPU commands: `" + TR + " bench.py --gpus N ...`.", ""]
blocks = [
    ("C3 (headline), 1 GPU: `python bench.py`", "f_c3.json"),
    ("C3, 2 GPUs, weak: `bench.py --gpus 2`", "g2_c3_weak.json"),
    ("C3, 2 GPUs, strong (4096 instances in total): `bench.py --gpus 2 --scaling strong`", "g2_c3_strong.json"),
    ("C3, 8 GPUs, weak: `bench.py --gpus 8`", "g8_c3_weak.json"),
    ("C3, 8 GPUs, strong (4096 instances in total): `bench.py --gpus 8 --scaling strong`", "g8_c3_strong.json"),
    ("C5 (1000 EVSEs, general path, 128 instances per GPU), 1 GPU: `python bench.py --config c5 --steps 4 --no-cpu-baseline --no-latency` (build before the 2-warp row blocks)", "t11_c5.json"),
    ("C5, 2 GPUs: `bench.py --gpus 2 --config c5 --steps 4`", "g2_c5.json"),
    ("C5, 8 GPUs: `bench.py --gpus 8 --config c5 --steps 4`", "g8_c5.json"),
    ("C1 (one instance per step): `python bench.py --config c1 --no-cpu-baseline --no-latency`", "h_c1.json"),
    ("C2 (one instance per step): `python bench.py --config c2 --no-cpu-baseline --no-latency`", "h_c2.json"),
    ("C4 (closed-loop replay on the device, 1024 sites, 40 control steps from 8 am), 1 GPU: `python bench.py --config c4 --steps 40 --no-cpu-baseline --no-latency`", "h_c4.json"),
    ("C4, 2 GPUs: `bench.py --gpus 2 --config c4 --steps 40`", "g2_c4.json"),
    ("C4, 8 GPUs: `bench.py --gpus 8 --config c4 --steps 40`", "g8_c4.json"),
]
summary = []
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
    summary.append((title.split(":")[0], l["n_gpus"], l["scaling"], l["value"], l["e2e"]["value"] if l.get("e2e") else None, l["ms_per_step"],
                    (l.get("roofline") or {}).get("frac")))
out += ["## Summary", "", "| run | GPUs | scaling | value (solves/s) | e2e (solves/s) | ms per step | roofline.frac |", "|---|---|---|---|---|---|---|"]
for t, n, sc, v, e, ms, fr in summary:
    out.append(f"| {t} | {n} | {sc} | {v:,.0f} | {e:,.0f} | {ms:.2f} | {'' if fr is None else f'{fr:.3f}'} |" if e is not None else f"| {t} | {n} | {sc} | {v:,.0f} | | {ms:.2f} | |")
l30 = load("h_c4_30d.json")
if l30:
    out += ["", "## Closed-loop replay, full BASELINE config 4: `python tools/replay_c4.py 1024 0 8640 30 128 device 8`", "", "```json", json.dumps(l30), "```", "",
            "30 days x 1024 sites x 288 steps = 5.07 M warm-started MPC solves in **10.0 s on one B200** (mid-round build: 11.1 s; round 1: 39.0 s)."]
out += ["", "GPU tests of the same build: `pytest tests -m gpu` 148 passed (58 s); `__graft_entry__.smoke()`: objective 987.068142 = oracle, "
        "max |dR| 2.2e-06 A, 50 iterations.", ""]
open(os.path.join(ROOT, "profiles", "r02_final_bench_results.md"), "w").write("\n".join(out))
print("\n".join(out[-20:]))
