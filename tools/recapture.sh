set -u
O=gpurun_out
S="--no-e2e --no-cpu-baseline --no-secondary"
python bench.py --steps 6 --warmup 3 $S > $O/r2_plain_10x10.json 2> $O/recap.err && \
ncu --set full --clock-control none --import-source on -k regex:te_step_kernel -s 103 -c 1 -f -o $O/r2_final_10x10 python bench.py --steps 6 --warmup 3 $S > $O/ncu_r2_final_10x10.log 2>&1
python bench.py --workload grid3x3_L250_greedy --steps 6 --warmup 3 $S > $O/r2_plain_3x3.json 2>> $O/recap.err && \
ncu --set full --clock-control none --import-source on -k regex:te_step_kernel -s 53 -c 1 -f -o $O/r2_final_3x3 python bench.py --workload grid3x3_L250_greedy --steps 6 --warmup 3 $S > $O/ncu_r2_final_3x3.log 2>&1
ls -la $O/r2_final_*.ncu-rep
