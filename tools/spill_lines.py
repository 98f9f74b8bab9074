"""DEV: where the register spills of one solve-kernel instantiation are: STL / LDL instructions per source line
(nvdisasm -g on the cubin extracted from the built library).  python tools/spill_lines.py [Q TPW MULTI FAST] [lib]"""
import collections
import os
import re
import subprocess
import sys
import tempfile

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
args = sys.argv[1:]
q, tpw, multi, fast = (args + ["9", "3", "0", "1"])[:4] if len(args) < 4 else args[:4]
lib = args[4] if len(args) > 4 else os.path.join(ROOT, "adacharge_b200", "libadacharge_b200.so")
name = f"_Z16acb_solve_kernelILi{q}ELi{tpw}ELb{multi}ELb{fast}EE"
with tempfile.TemporaryDirectory() as d:
    subprocess.run(["cuobjdump", "-xelf", "all", lib], cwd=d, check=True, stdout=subprocess.DEVNULL)
    cub = os.path.join(d, f"acb_solve_q{q}.sm_100a.cubin")
    sass = subprocess.run(["nvdisasm", "-g", cub], capture_output=True, text=True).stdout.splitlines()
inside, cur = False, None
st, ld, n = collections.Counter(), collections.Counter(), 0
for line in sass:
    if line.startswith(".text."):
        inside = name in line
        continue
    if not inside:
        continue
    m = re.search(r'//## File "(.*?)", line (\d+)', line)
    if m:
        cur = (os.path.basename(m.group(1)), int(m.group(2)))
        continue
    if re.match(r"\s+/\*[0-9a-f]+\*/", line):
        n += 1
        if re.search(r"\bSTL\b", line):
            st[cur] += 1
        if re.search(r"\bLDL\b", line):
            ld[cur] += 1
print(f"{name}: {n} instructions, {sum(st.values())} STL, {sum(ld.values())} LDL")
for k in sorted(set(st) | set(ld), key=lambda x: (x[0], x[1])):
    print(f"{k[0]}:{k[1]:5d}  STL {st[k]:3d}  LDL {ld[k]:3d}")
