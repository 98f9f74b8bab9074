"""DEV: per-source-line instruction counts and stall samples of one kernel from an ncu report.
The SASS page of the report (ncu --page source --csv) is joined, instruction by instruction, with the line table of the
built library (nvdisasm -g on the extracted cubin).  python tools/ncu_lines.py <report.ncu-rep> <kernel-regex> <cubin-stem> <mangled-substring> [top]"""
import collections, csv, io, os, re, subprocess, sys, tempfile

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
rep, kre, stem, mangled = sys.argv[1:5]
top = int(sys.argv[5]) if len(sys.argv) > 5 else 40
lib = os.path.join(ROOT, "adacharge_b200", "libadacharge_b200.so")
txt = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--kernel-name", f"regex:{kre}"], capture_output=True, text=True).stdout
lines = txt.splitlines()
h = next(i for i, l in enumerate(lines) if l.startswith('"Address"'))
end = next((i for i in range(h + 1, len(lines)) if lines[i].startswith('"Kernel Name"')), len(lines))
rows = list(csv.DictReader(io.StringIO("\n".join(lines[h:end]))))
with tempfile.TemporaryDirectory() as d:
    subprocess.run(["cuobjdump", "-xelf", "all", lib], cwd=d, check=True, stdout=subprocess.DEVNULL)
    sass = subprocess.run(["nvdisasm", "-g", os.path.join(d, f"{stem}.sm_100a.cubin")], capture_output=True, text=True).stdout.splitlines()
inside, cur, tab = False, None, []
for l in sass:
    if l.startswith(".text."):
        inside = mangled in l
        continue
    if not inside:
        continue
    m = re.search(r'//## File "(.*?)", line (\d+)', l)
    if m:
        cur = (m.group(1), int(m.group(2)))
        continue
    if re.match(r"\s+/\*[0-9a-f]+\*/", l):
        tab.append(cur)
assert len(tab) >= len(rows), (len(tab), len(rows))
_src = {}
def text(key):
    if not key:
        return ""
    f, ln = key
    if f not in _src:
        try:
            _src[f] = open(f if os.path.isabs(f) else os.path.join(ROOT, "adacharge_b200", "csrc", f)).read().splitlines()
        except OSError:
            _src[f] = []
    return _src[f][ln - 1].strip()[:120] if 0 < ln <= len(_src[f]) else ""
ins, smp = collections.Counter(), collections.Counter()
for r, ln in zip(rows, tab):
    ins[ln] += int(r["Instructions Executed"] or 0)
    smp[ln] += int(r["# Samples"] or 0)
ti, ts = sum(ins.values()), sum(smp.values())
print(f"{len(rows)} SASS instructions, {ti} executed, {ts} samples")
for ln, _ in smp.most_common(top):
    print(f"{100 * smp[ln] / ts:5.1f}% samples {100 * ins[ln] / ti:5.1f}% inst  {os.path.basename(ln[0]) if ln else '?'}:{ln[1] if ln else 0}: {text(ln)}")
