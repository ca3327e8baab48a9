(time timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "early_copy_out or config1_full or back_to_back" 2>&1 | tail -4) > gpurun_out/r3o_pytest.txt 2>&1
tail -3 gpurun_out/r3o_pytest.txt
(time timeout 600 python bench.py --steps 5 --warmup 3 --no-host-shim) > gpurun_out/r3o_bench.json 2> gpurun_out/r3o_bench.err
python - <<'PY'
import json
d=json.loads(open("gpurun_out/r3o_bench.json").read().strip().splitlines()[-1])
print(d["ms_per_step"], d["e2e"]["value"], d["e2e"]["stages_ms_last_step"], d["result"]["bit_exact_vs_cpu_port"])
PY
echo done
