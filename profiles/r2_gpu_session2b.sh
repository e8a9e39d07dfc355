set -x
for wl in c1 c2; do
  ncu --set full --clock-control none --import-source on -k regex:self_ring_vec_kernel -s 152 -c 1 -o gpurun_out/r2_vec_$wl python bench.py --workload $wl --steps 5 --warmup 3 --no-cpu-baseline --no-other-workloads --no-parity-check --no-e2e > gpurun_out/r2_ncu_vec_$wl.log 2>&1; echo $wl rc=$?
done
