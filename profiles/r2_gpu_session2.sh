# small shapes: sensitivity of the vector kernel to tile scheduling, and one full ncu capture each
set -x
F="--steps 50 --warmup 5 --no-cpu-baseline --no-other-workloads --no-parity-check --no-e2e"
for wl in c1 c2; do
  for env in "X=1" "SQT_DETERMINISTIC=1" "SQT_TILE=64" "SQT_TILE=128" "SQT_TILE=32" "SQT_FUSE=0"; do
    env $env python bench.py --workload $wl $F 2>/dev/null | python -c "
import sys,json
d=json.loads([l for l in sys.stdin if l.startswith('{')][-1])
print('$wl','$env',d['ms_per_step'],d['roofline']['avg_launch_ms'],d['gpu_launches'],d['config']['kernel'][:110])"
  done
done
for wl in c1 c2; do
  ncu --set full --clock-control none --import-source on -k regex:self_ring_vec_kernel -s 200 -c 1 -o gpurun_out/r2_vec_$wl python bench.py --workload $wl --steps 5 --warmup 3 --no-cpu-baseline --no-other-workloads --no-parity-check --no-e2e > gpurun_out/r2_ncu_vec_$wl.log 2>&1; echo $wl rc=$?
done
