set -x
A="python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-other-workloads --no-parity-check --no-e2e"
python bench.py --steps 10 --warmup 3 > gpurun_out/r2_bench_c5_1gpu.json 2> gpurun_out/r2_bench_c5_1gpu.err; echo bench rc=$?
$A > gpurun_out/r2_plainA.log 2>&1 && ncu --metrics gpu__time_duration.sum --clock-control none -s 880 -c 100 --csv --log-file gpurun_out/r2_ncu_launches_c5.csv $A > gpurun_out/r2_ncu_launch_run.log 2>&1; echo launches rc=$?
ncu --set full --clock-control none --import-source on -k regex:self_ring_mma_kernel -s 152 -c 1 -o gpurun_out/r2_mma_c5 $A > gpurun_out/r2_ncu_full_run.log 2>&1; echo full rc=$?
C3="python bench.py --workload c3 --steps 2 --warmup 3 --no-cpu-baseline --no-other-workloads --no-parity-check --no-e2e"
ncu --set full --clock-control none --import-source on -k regex:self_ring_mma_kernel -s 152 -c 1 -o gpurun_out/r2_mma_c3 $C3 > gpurun_out/r2_ncu_full_c3_run.log 2>&1; echo c3 rc=$?
C1="python bench.py --workload c1 --steps 20 --warmup 3 --no-cpu-baseline --no-other-workloads --no-parity-check --no-e2e"
ncu --metrics gpu__time_duration.sum --clock-control none -s 320 -c 60 --csv --log-file gpurun_out/r2_ncu_launches_c1.csv $C1 > gpurun_out/r2_ncu_launch_c1_run.log 2>&1; echo c1 rc=$?
tail -c 600 gpurun_out/r2_bench_c5_1gpu.json
