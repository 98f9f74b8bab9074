"""Turn ncu CSV output into the markdown summaries kept under profiles/.
  python tools/ncu_summarize.py launches <launches.csv> <title>      -> launch list table on stdout
  python tools/ncu_summarize.py raw <raw.csv> <kernel-regex>          -> selected metrics of the first matching kernel"""
import csv, io, re, sys

def rows(path):
    txt = [l for l in open(path, errors="replace") if l.startswith('"')]
    return list(csv.DictReader(io.StringIO("".join(txt))))

def launches(path, title):
    per = {}
    for r in rows(path):
        d = per.setdefault(int(r["ID"]), dict(k=r["Kernel Name"], grid=r["Grid Size"], block=r["Block Size"]))
        d[r["Metric Name"]] = float(r["Metric Value"].replace(",", "")) * ({"msecond": 1e6, "usecond": 1e3, "second": 1e9}.get(r["Metric Unit"], 1) if r["Metric Name"].startswith("gpu__time") else 1)
    print(f"# {title}\n")
    print("| # | kernel | grid | block | time (ms) | dram read (MB) | dram write (MB) |\n|---|---|---|---|---|---|---|")
    tot, mine, dr = 0.0, 0.0, []
    for i in sorted(per):
        d = per[i]
        t = d.get("gpu__time_duration.sum", 0) / 1e6
        tot += t
        if "acb_" in d["k"]:
            mine += t
            dr.append(d.get("dram__bytes_read.sum", 0) + d.get("dram__bytes_write.sum", 0))
        print(f"| {i} | `{d['k'][:72]}` | {d['grid']} | {d['block']} | {t:.3f} | {d.get('dram__bytes_read.sum', 0) / 1e6:.1f} | {d.get('dram__bytes_write.sum', 0) / 1e6:.1f} |")
    print(f"\nShare of kernel time in this library's kernels: {100 * mine / max(tot, 1e-12):.1f} %.")
    if dr:
        print(f"Mean DRAM traffic per acb launch: {sum(dr) / len(dr) / 1e6:.1f} MB.")

KEEP = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
        "smsp__inst_executed.sum", "sm__throughput.avg.pct_of_peak_sustained_elapsed", "smsp__issue_active.avg.pct_of_peak_sustained_active",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "launch__registers_per_thread", "launch__block_size", "launch__grid_size",
        "launch__shared_mem_per_block_dynamic", "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active",
        "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum.pct_of_peak_sustained_elapsed", "sass__inst_executed_local_loads", "sass__inst_executed_local_stores",
        "smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio", "smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio", "smsp__average_warps_issue_stalled_wait_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_no_instruction_per_issue_active.ratio", "smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_mio_throttle_per_issue_active.ratio", "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum"]

def raw(path, rx):
    rs = rows(path)
    units = rs[0]
    for r in rs[1:]:
        if re.search(rx, r.get("Kernel Name", "")):
            print(f"kernel: `{r['Kernel Name'][:100]}`  grid {r.get('Grid Size')} block {r.get('Block Size')}\n")
            print("| metric | value |\n|---|---|")
            for k in KEEP:
                if k in r:
                    print(f"| `{k}` ({units.get(k, '')}) | {r[k]} |")
            return
    print("no kernel matched")

if __name__ == "__main__":
    {"launches": launches, "raw": raw}[sys.argv[1]](sys.argv[2], sys.argv[3])
