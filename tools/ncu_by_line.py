"""Join an ncu SASS source page (per-instruction counters) with nvdisasm line info -> hot source lines.

usage: python tools/ncu_by_line.py <report.ncu-rep> <kernel regex> [top N]
"""
import collections
import csv
import io
import re
import subprocess
import sys
import os

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
rep, kre = sys.argv[1], sys.argv[2]
top = int(sys.argv[3]) if len(sys.argv) > 3 else 40
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--kernel-name", f"regex:{kre}"], capture_output=True, text=True).stdout
# the report may hold several launches of the kernel: take the first block
blocks = out.split('"Kernel Name"')
blk = '"Kernel Name"' + blocks[1]
rows = list(csv.reader(io.StringIO(blk)))
kname = rows[0][1]
h = rows[1]
si, ie, ss = h.index("Source"), h.index("Instructions Executed"), h.index("# Samples")
sass = [(r[si].strip(), int(r[ie]), int(r[ss]) if r[ss].isdigit() else 0) for r in rows[2:] if len(r) > ie and r[ie].isdigit()]
# line info from the cubin
so = os.environ.get("NCU_SO") or os.path.join(ROOT, "pyft8_b200", "libft8_b200.so")
tmp = "/tmp/_cubin"
os.makedirs(tmp, exist_ok=True)
subprocess.run(["cuobjdump", "-xelf", "all", so], cwd=tmp, capture_output=True)
cubin = [os.path.join(tmp, f) for f in os.listdir(tmp) if f.endswith(".cubin")][0]
mangled = re.match(r"(?:void )?(?:ft8::)?(\w+)", kname).group(1)
dis = subprocess.run(["nvdisasm", "-g", "-c", cubin], capture_output=True, text=True).stdout
lines, cur, infn = [], None, False
for l in dis.splitlines():
    if l.startswith(".text."):
        infn = mangled in l
        continue
    if not infn:
        continue
    m = re.search(r'//## File "([^"]+)", line (\d+)', l)
    if m:
        cur = (os.path.basename(m.group(1)), int(m.group(2)))
        continue
    if re.match(r"\s+/\*[0-9a-f]{4,6}\*/", l):
        lines.append(cur)
n = min(len(lines), len(sass))
agg_i, agg_s = collections.Counter(), collections.Counter()
for i in range(n):
    agg_i[lines[i]] += sass[i][1]
    agg_s[lines[i]] += sass[i][2]
ti, ts = sum(agg_i.values()), sum(agg_s.values())
print(f"{kname[:80]}: {len(sass)} SASS instrs, {len(lines)} with line info; total warp-instr {ti}, samples {ts}")
src_cache = {}
by_inst = os.environ.get("BY_INST") == "1"       # BY_INST=1: rank lines by executed instructions instead of stall samples
for key, v in sorted((agg_i if by_inst else agg_s).items(), key=lambda x: -x[1])[:top]:
    v = agg_s[key]
    txt = ""
    if key:
        for d in ("pyft8_b200/csrc", "include"):
            pth = os.path.join(ROOT, d, key[0])
            if os.path.exists(pth):
                src_cache.setdefault(pth, open(pth).read().splitlines())
                txt = src_cache[pth][key[1] - 1].strip()[:90]
    print(f"{v/ts*100:5.1f}% smp {agg_i[key]/ti*100:5.1f}% inst  {str(key):28s} {txt}")
