"""Developer tool: share of executed instructions / stall samples / active lanes per REGION of te_step_kernel.

Joins an ncu source page (csv) with `nvdisasm -gi` (line info WITH the inlining chain), attributes every SASS
instruction to the outermost te_kernels.cuh line of the kernel body it was inlined into, and sums over line ranges.

  ncu -i X.ncu-rep --page source --csv > src.csv
  cuobjdump -xelf all traffic_env_b200/libtraffic_b200.so ; nvdisasm -gi -c *.cubin > disi.txt
  python tools/ncu_regions.py src.csv disi.txt <mangled-kernel-substring> regions.txt [kernel-name-substring-in-report]

regions.txt holds a Python dict {name: (first_line, last_line)} for te_kernels.cuh.
"""
import csv
import re
import sys
from collections import defaultdict

src_csv, dis_txt, dis_kern, regions_file = sys.argv[1:5]
rep_kern = sys.argv[5] if len(sys.argv) > 5 else "te_step"
regions = eval(open(regions_file).read())

in_k = False
chain = []
pending_new = True
dis = []
for ln in open(dis_txt):
    if ln.startswith("\t.section\t.text."):
        in_k = dis_kern in ln
        continue
    if not in_k:
        continue
    m = re.search(r'//## File "([^"]+)", line (\d+)', ln)
    if m:
        if pending_new:
            chain = []
            pending_new = False
        chain.append((m.group(1).split("/")[-1], int(m.group(2))))
        continue
    m = re.match(r"\s+/\*([0-9a-f]{4,})\*/\s+(.*?);", ln)
    if m:
        dis.append((m.group(2).strip(), list(chain)))
        pending_new = True

rows = list(csv.reader(open(src_csv)))
start = [i for i, r in enumerate(rows) if len(r) >= 2 and r[0] == "Kernel Name" and rep_kern in r[1]][0]
hdr = rows[start + 1]
iI, iT, iS = hdr.index("Instructions Executed"), hdr.index("Thread Instructions Executed"), hdr.index("# Samples")
body = []
for r in rows[start + 2:]:
    if len(r) < len(hdr) or r[0] == "Kernel Name":
        break
    body.append(r)
assert len(body) == len(dis), (len(body), len(dis))


def region_of(ch):
    outer = None
    for f, l in ch:          # innermost first; the outermost kernel-body line is the last te_kernels.cuh entry
        if f == "te_kernels.cuh":
            outer = l
    if outer is None:
        return "other"
    for name, (a, b) in regions.items():
        if a <= outer <= b:
            return name
    return "line %d" % outer


per = defaultdict(lambda: [0, 0, 0, 0])
ti = tt = ts = 0
for r, (sass, ch) in zip(body, dis):
    ii, t, s = int(r[iI]), int(r[iT]), int(r[iS])
    key = region_of(ch)
    is_math = 1 if (ch and ch[0][0] == "te_math.cuh") else 0
    per[key][0] += ii; per[key][1] += t; per[key][2] += s; per[key][3] += ii * is_math
    ti += ii; tt += t; ts += s
print("total warp-instructions %.4g, active lanes per instruction %.1f, stall samples %d" % (ti, tt / ti, ts))
print("%-24s %7s %7s %6s %9s" % ("region", "inst%", "smp%", "lanes", "te_math%"))
for k, (ii, t, s, mth) in sorted(per.items(), key=lambda kv: -kv[1][0]):
    print("%-24s %7.2f %7.2f %6.1f %9.2f" % (k, 100.0 * ii / ti, 100.0 * s / max(ts, 1), t / max(ii, 1), 100.0 * mth / ti))
