# full-size BASELINE config 4 on one GPU: 8 shard aggregators merged in process vs one aggregator over the same stream
timeout 300 python bench.py --config4-local 8 --merge-rows 200000 --steps 2 --warmup 1 --e2e-steps 2 > gpurun_out/r2j_c4_local_small.json 2> gpurun_out/r2j_c4_local_small.err
timeout 300 python bench.py --config4-single 8 --merge-rows 200000 --steps 2 --warmup 1 --e2e-steps 2 > gpurun_out/r2j_c4_single_small.json 2> gpurun_out/r2j_c4_single_small.err
python - <<'PY'
import json
a=json.loads(open("gpurun_out/r2j_c4_local_small.json").read().strip().splitlines()[-1]); b=json.loads(open("gpurun_out/r2j_c4_single_small.json").read().strip().splitlines()[-1])
print("small digests equal:", a["ipc_sha256"]==b["ipc_sha256"], a["rows"], b["rows"], a["ms_per_step"], b["ms_per_step"])
PY
free -g | head -2
(time timeout 1100 python bench.py --config4-local 8 --steps 3 --warmup 1 --e2e-steps 2) > gpurun_out/r2j_c4_local.json 2> gpurun_out/r2j_c4_local.err
tail -c 300 gpurun_out/r2j_c4_local.err
(time timeout 1100 python bench.py --config4-single 8 --steps 3 --warmup 1 --e2e-steps 2) > gpurun_out/r2j_c4_single.json 2> gpurun_out/r2j_c4_single.err
tail -c 300 gpurun_out/r2j_c4_single.err
python - <<'PY'
import json
try:
    a=json.loads(open("gpurun_out/r2j_c4_local.json").read().strip().splitlines()[-1]); b=json.loads(open("gpurun_out/r2j_c4_single.json").read().strip().splitlines()[-1])
    print("full digests equal:", a["ipc_sha256"]==b["ipc_sha256"], a["rows"], b["rows"], a["ms_per_step"], b["ms_per_step"])
except Exception as e:
    print("ERR", e)
PY
echo done
