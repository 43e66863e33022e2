#!/bin/bash
# Build-container side of an ncu capture: raw metrics (selected), per-region and per-source-line breakdowns.
#   tools/ncu_summarise.sh <capture.ncu-rep> <mangled-kernel-substring> <out-prefix under profiles/>
set -e
REP=$1; KERN=$2; OUT=$3
T=$(mktemp -d)
( cd $T && cuobjdump -xelf all /root/repo/traffic_env_b200/libtraffic_b200.so > /dev/null && nvdisasm -c -gi te_api.sm_100a.cubin > disi.txt && nvdisasm --print-line-info -c te_api.sm_100a.cubin > dis.txt )
ncu -i $REP --page source --csv > $T/src.csv 2>/dev/null
ncu -i $REP --page raw --csv > $T/raw.csv 2>/dev/null
python - "$T/raw.csv" > profiles/${OUT}_metrics.csv <<'PY'
import csv, sys
rows = list(csv.reader(open(sys.argv[1])))
h, u, v = rows[0], rows[1], rows[2]
keep = ("Kernel Name", "gpu__time_duration.sum", "smsp__inst_executed.sum", "smsp__thread_inst_executed_per_inst_executed.ratio",
        "smsp__issue_active.avg.pct_of_peak_sustained_active", "sm__warps_active.avg.pct_of_peak_sustained_active",
        "dram__bytes_read.sum", "dram__bytes_write.sum", "dram__throughput.avg.pct_of_peak_sustained_elapsed",
        "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
        "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum",
        "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum.pct_of_peak_sustained_elapsed",
        "launch__registers_per_thread", "launch__block_size", "launch__grid_size", "launch__shared_mem_per_block_dynamic",
        "launch__occupancy_limit_registers", "launch__occupancy_limit_shared_mem", "launch__occupancy_limit_warps",
        "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
        "smsp__sass_inst_executed_op_local_ld.sum", "smsp__sass_inst_executed_op_local_st.sum")
w = csv.writer(sys.stdout)
w.writerow(["metric", "unit", "value"])
for a, b, c in zip(h, u, v):
    if a in keep or a.startswith("smsp__average_warps_issue_stalled_") and a.endswith("_per_issue_active.ratio"):
        w.writerow([a, b, c])
PY
python - > $T/regions.txt <<'PY'
import re
src = open("/root/repo/traffic_env_b200/csrc/te_kernels.cuh").read().splitlines()
def find(pat, start=0):
    for i in range(start, len(src)):
        if pat in src[i]:
            return i + 1
    raise SystemExit("pattern not found: " + pat)
k = find("__global__ void __launch_bounds__(MAXT, MINB) te_step_kernel")
marks = [("prologue_stage", k), ("prologue_arrivals", find("for (int g = warp; g < ng; g += nwarps)")),
         ("prologue_lights_init", find("int li_ph = 0")), ("prologue_sort", find("row -> (warp, lane) assignment")),
         ("prologue_regs", find("lane = road: ring indices, counters")), ("tick_arrivals", find("for (int step = 0;; step++)")),
         ("tick_lights", find("{  // update_lights")), ("tick_list_setup", find("const int n = frozen ? 0")),
         ("car_loop", find("unsigned int acc = 0u;")), ("tick_pops", find("int npop = 0;")),
         ("phaseC", find("phase C", find("int npop = 0;"))), ("step_epilogue", find("end of the actor step")),
         ("flush", find("---- flush")), ("end", find("// Staging only"))]
print("{" + ", ".join('"%s": (%d, %d)' % (marks[i][0], marks[i][1], marks[i + 1][1] - 1) for i in range(len(marks) - 1)) + "}")
PY
python tools/ncu_regions.py $T/src.csv $T/disi.txt $KERN $T/regions.txt > profiles/${OUT}_regions.txt
DIS_KERNEL=$KERN python tools/ncu_lines.py $T/src.csv $T/dis.txt te_step_kernel 40 > profiles/${OUT}_source_lines.txt
head -16 profiles/${OUT}_regions.txt
rm -rf $T
