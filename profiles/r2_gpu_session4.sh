F="--steps 50 --warmup 5 --no-cpu-baseline --no-other-workloads --no-parity-check --no-e2e"
run() { wl=$1; shift; env "$@" python bench.py --workload $wl $F 2>/dev/null | python -c "
import sys,json
d=json.loads([l for l in sys.stdin if l.startswith('{')][-1])
print('$wl','$*','ms/frame',round(d['ms_per_step']*1e3,2),'us  kernel',round(d['roofline']['avg_launch_ms']*1e3,2),'us', d['config']['kernel'].split(' persistent')[1][:60])"; }
run c1 X=1
run c1 SQT_VEC_MINB=3
run c1 SQT_VEC_MINB=3 SQT_DETERMINISTIC=1
run c2 X=1
run c2 SQT_VEC_MINB=5
run c2 SQT_VEC_SPW=1 SQT_VEC_MINB=6
run c2 SQT_VEC_SPW=1 SQT_VEC_MINB=8
run c2 SQT_VEC_SPW=1 SQT_VEC_MINB=6 SQT_TILE=64
run c2 SQT_VEC_SPW=1 SQT_VEC_MINB=6 SQT_TILE=128
run c2 SQT_TILE=64
