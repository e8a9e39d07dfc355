# 8 GPUs of one box: the bench line at N = 8 and N = 2 (parity_check of the sharded path included) and the ingest-included drop-in run
for N in 8 2; do
python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 2953$N bench.py --gpus $N --steps 20 --warmup 3 > gpurun_out/r2_scale_$N.json 2> gpurun_out/r2_scale_$N.err; echo bench$N rc=$?
done
python profiles/cli_ingest_c3.py 240 1,8 > gpurun_out/r2_cli_ingest_c3.json 2> gpurun_out/r2_cli_ingest_c3.err; echo ingest rc=$?
tail -c 1800 gpurun_out/r2_cli_ingest_c3.json; tail -3 gpurun_out/r2_cli_ingest_c3.err
