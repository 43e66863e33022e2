"""Developer tool: join an ncu SASS-level source page (csv) with nvdisasm line info and print the
instruction / stall-sample share of every CUDA source line of te_step_kernel.

  ncu -i X.ncu-rep --page source --csv > src.csv
  cuobjdump -xelf all traffic_env_b200/libtraffic_b200.so ; nvdisasm --print-line-info -c *.cubin > dis.txt
  python tools/ncu_lines.py src.csv dis.txt [kernel_substring]
"""
import csv
import re
import sys
from collections import defaultdict

src_csv, dis_txt = sys.argv[1], sys.argv[2]
kern = sys.argv[3] if len(sys.argv) > 3 else "te_step_kernel"
import os
dis_kern = os.environ.get("DIS_KERNEL", kern)  # mangled-name substring for the disassembly (templates)

# --- nvdisasm: instructions of the kernel in order, with the current //## File "...", line N marker
lines = open(dis_txt).read().splitlines()
in_k = False
cur = ("?", 0)
inline_stack = ""
dis = []
for ln in lines:
    if ln.startswith("\t.section\t.text."):
        in_k = dis_kern in ln
        continue
    if not in_k:
        continue
    m = re.search(r'//## File "([^"]+)", line (\d+)(.*)', ln)
    if m:
        cur = (m.group(1).split("/")[-1], int(m.group(2)))
        inline_stack = m.group(3)
        continue
    m = re.match(r"\s+/\*([0-9a-f]{4,})\*/\s+(.*?);", ln)
    if m:
        dis.append((int(m.group(1), 16), m.group(2).strip(), cur, inline_stack))

rows = list(csv.reader(open(src_csv)))
# find the kernel block
start = None
for i, r in enumerate(rows):
    if len(r) >= 2 and r[0] == "Kernel Name" and kern in r[1]:
        start = i
        break
hdr = rows[start + 1]
iI, iT, iS = hdr.index("Instructions Executed"), hdr.index("Thread Instructions Executed"), hdr.index("# Samples")
body = []
for r in rows[start + 2:]:
    if len(r) < len(hdr) or r[0] == "Kernel Name":
        break
    body.append(r)
print("sass rows in report %d, in disassembly %d" % (len(body), len(dis)))
n = min(len(body), len(dis))
per = defaultdict(lambda: [0, 0, 0])
tot_i = tot_s = tot_t = 0
for k in range(n):
    r = body[k]
    ii, tt, ss = int(r[iI]), int(r[iT]), int(r[iS])
    key = dis[k][2]
    per[key][0] += ii; per[key][1] += tt; per[key][2] += ss
    tot_i += ii; tot_t += tt; tot_s += ss
print("total warp-instructions %d, thread-instructions %d (%.1f thr/inst), samples %d" % (tot_i, tot_t, tot_t / tot_i, tot_s))
srcs = {}
def srcline(f, l):
    if f not in srcs:
        try:
            import glob
            p = glob.glob("/root/repo/traffic_env_b200/csrc/" + f) or glob.glob("/root/repo/**/" + f, recursive=True)
            srcs[f] = open(p[0]).read().splitlines()
        except Exception:
            srcs[f] = []
    return srcs[f][l - 1].strip()[:100] if 0 < l <= len(srcs[f]) else ""
out = sorted(per.items(), key=lambda kv: -kv[1][0])
print("%6s %6s %5s  %s" % ("inst%", "smp%", "thr", "line"))
for (f, l), (ii, tt, ss) in out[:int(sys.argv[4]) if len(sys.argv) > 4 else 45]:
    print("%6.2f %6.2f %5.1f  %s:%d  %s" % (100 * ii / tot_i, 100 * ss / max(tot_s, 1), tt / max(ii, 1), f, l, srcline(f, l)))

if len(sys.argv) > 5:
    # dump SASS rows attributed to FILE:LINE with their execution counts
    f, l = sys.argv[5].split(":")
    prev = None
    for k in range(n):
        if dis[k][2] == (f, int(l)):
            r = body[k]
            if prev is not None and k != prev + 1:
                print("   ...")
            print("%6d %10s inst %5.1f thr  %s   %s" % (k, r[iI], int(r[iT]) / max(int(r[iI]), 1), dis[k][1][:70], dis[k][3][:60]))
            prev = k
