#!/bin/bash
# Round-end evidence on one B200 (run under gpurun): tests, bench lines, launch lists, full ncu captures.
# Every ncu run comes directly after the same command line has exited 0 without ncu.
set -u
O=gpurun_out
B="python bench.py --steps 21 --warmup 6"
S="--no-e2e --no-cpu-baseline --no-secondary"
python -m pytest tests -m gpu -q > $O/r2_final_gputests.log 2>&1; tail -3 $O/r2_final_gputests.log
python -c "import __graft_entry__ as g; g.smoke()" > $O/r2_final_smoke.log 2>&1; tail -1 $O/r2_final_smoke.log
$B > $O/bench_r2_final_n1.json 2> $O/bench_r2_final_n1.err
$B --workload grid3x3_L250_greedy --no-secondary > $O/bench_r2_final_3x3_n1.json 2>> $O/bench_r2_final_n1.err
$B --no-multi $S > $O/bench_r2_final_n1_nomulti.json 2>> $O/bench_r2_final_n1.err
$B --no-multi $S --workload grid3x3_L250_greedy > $O/bench_r2_final_3x3_n1_nomulti.json 2>> $O/bench_r2_final_n1.err
$B --launch-steps 3 $S > $O/bench_r2_final_n1_ls3.json 2>> $O/bench_r2_final_n1.err
$B --launch-steps 3 $S --workload grid3x3_L250_greedy > $O/bench_r2_final_3x3_n1_ls3.json 2>> $O/bench_r2_final_n1.err
# launch lists (the launches around and inside the timed region with their device times).  A launch is 6 actor steps (two
# greedy decisions): the 300-step (150-step) pre-roll is 50 (25) launches, then 1 warm-up launch, 2 timed launches, and the
# 10 launches of the per-launch timing pass - the full capture takes one of those.
python bench.py --steps 12 --warmup 6 $S > $O/r2_plain_10x10.json 2>> $O/bench_r2_final_n1.err && \
ncu --metrics gpu__time_duration.sum --clock-control none -s 48 -c 18 --csv --log-file $O/r2_final_launches_10x10.csv python bench.py --steps 12 --warmup 6 $S > /dev/null 2>&1
python bench.py --steps 12 --warmup 6 $S > $O/r2_plain_10x10.json 2>> $O/bench_r2_final_n1.err && \
ncu --set full --clock-control none --import-source on -k regex:te_step_kernel -s 58 -c 1 -f -o $O/r2_final_10x10 python bench.py --steps 12 --warmup 6 $S > $O/ncu_r2_final_10x10.log 2>&1
python bench.py --workload grid3x3_L250_greedy --steps 12 --warmup 6 $S > $O/r2_plain_3x3.json 2>> $O/bench_r2_final_n1.err && \
ncu --metrics gpu__time_duration.sum --clock-control none -s 23 -c 18 --csv --log-file $O/r2_final_launches_3x3.csv python bench.py --workload grid3x3_L250_greedy --steps 12 --warmup 6 $S > /dev/null 2>&1
python bench.py --workload grid3x3_L250_greedy --steps 12 --warmup 6 $S > $O/r2_plain_3x3.json 2>> $O/bench_r2_final_n1.err && \
ncu --set full --clock-control none --import-source on -k regex:te_step_kernel -s 33 -c 1 -f -o $O/r2_final_3x3 python bench.py --workload grid3x3_L250_greedy --steps 12 --warmup 6 $S > $O/ncu_r2_final_3x3.log 2>&1
python tests/big_soak_run.py > $O/r2_final_soak.log 2>&1; tail -8 $O/r2_final_soak.log
python tools/peak_latency.py > $O/r2_peak_latency.txt 2>&1; tail -3 $O/r2_peak_latency.txt
python bench.py --impl reference --steps 20 --warmup 5 > $O/bench_r2_final_reference.json 2>> $O/bench_r2_final_n1.err
ls -la $O | grep r2_final
