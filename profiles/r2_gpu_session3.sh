F="--steps 50 --warmup 5 --no-cpu-baseline --no-other-workloads --no-parity-check --no-e2e"
for wl in c1 c2; do
  for env in "X=1" "SQT_DETERMINISTIC=1" "SQT_TILE=64" "SQT_TILE=128" $EXTRA_ENVS; do
    env $env python bench.py --workload $wl $F 2>/dev/null | python -c "
import sys,json
d=json.loads([l for l in sys.stdin if l.startswith('{')][-1])
print('$wl','$env','ms/frame',round(d['ms_per_step']*1e3,2),'us  kernel',round(d['roofline']['avg_launch_ms']*1e3,2),'us', d['config']['kernel'].split(' persistent')[1][:60])"
  done
done
