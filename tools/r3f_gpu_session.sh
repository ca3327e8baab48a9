(time timeout 600 python -m pytest tests/test_cpp_reporter.py tests/test_padata.py -m gpu -x -q 2>&1 | tail -5) > gpurun_out/r3f_pytest.txt 2>&1
tail -4 gpurun_out/r3f_pytest.txt
(time timeout 600 python bench.py) > gpurun_out/r3f_bench_default.json 2> gpurun_out/r3f_bench_default.err
python - <<'PY'
import json
d=json.loads(open("gpurun_out/r3f_bench_default.json").read().strip().splitlines()[-1])
print(d["ms_per_step"], d["e2e"]["value"], d["u32_ring"]["e2e"]["value"], d["result"]["bit_exact_vs_cpu_port"], d["cpu_baseline"]["value"])
print(json.dumps(d["host_shim"])[:1200])
PY
echo done
