# 8 GPUs: ingest-included drop-in runs at config-3 and config-4 size (raw binary frame files in /dev/shm)
df -h /dev/shm | tail -1; free -g | head -2
python profiles/cli_ingest_c3.py 240 1,8 > gpurun_out/r2_cli_ingest_c3.json 2> gpurun_out/r2_cli_ingest_c3.err; echo c3 rc=$?
python profiles/cli_ingest_c3.py 170 1,8 128 150 150 100 > gpurun_out/r2_cli_ingest_c4.json 2> gpurun_out/r2_cli_ingest_c4.err; echo c4 rc=$?
python - <<'P'
import json
for c in ("c3","c4"):
    try:
        d=json.load(open("gpurun_out/r2_cli_ingest_%s.json"%c))
        print(c, d["atoms"], d["frames"], round(d["file_GB"],1), "GB", {k:(round(v["ms_per_frame"],2), round(v["startup_s"],1), v["same_output_as_first_run"]) for k,v in d["runs"].items()})
    except Exception as ex: print(c, "failed", ex)
P
tail -2 gpurun_out/r2_cli_ingest_c4.err
