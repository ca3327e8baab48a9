(time timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "early_copy_out or config2_full or back_to_back or repeated" 2>&1 | tail -6) > gpurun_out/r3b_pytest.txt 2>&1
tail -5 gpurun_out/r3b_pytest.txt
PA_COLLECT_PROFILE=1 timeout 500 python bench.py --steps 5 --warmup 3 --no-host-shim > gpurun_out/r3b_bench.json 2> gpurun_out/r3b_bench.err
grep "collect:" gpurun_out/r3b_bench.err | tail -3
python - <<'PY'
import json
d=json.loads(open("gpurun_out/r3b_bench.json").read().strip().splitlines()[-1])
print(d["ms_per_step"], d["e2e"], d["u32_ring"]["e2e"], d["result"].get("bit_exact_vs_cpu_port"))
PY
echo done
