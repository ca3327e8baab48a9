(time timeout 900 python -m pytest tests/test_merge.py -m gpu -x -q -k "multi_process" 2>&1 | tail -6) > gpurun_out/r3e_pytest.txt 2>&1
tail -5 gpurun_out/r3e_pytest.txt
PA_ONE_GPU=1 timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29719 bench.py --gpus 2 --steps 5 --warmup 3 --merge-rows 4000000 --no-cpu > gpurun_out/r3e_dry_2proc.json 2> gpurun_out/r3e_dry_2proc.err
python - <<'PY'
import json
d=json.loads(open("gpurun_out/r3e_dry_2proc.json").read().strip().splitlines()[-1])
mb=d.get("mode_b") or {}
print(d["value"], mb.get("ms_per_step"), mb.get("error"), mb.get("exchange_payload_bytes_per_row"), mb.get("ipc_sha256"))
PY
echo done
