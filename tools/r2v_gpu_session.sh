(time timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "stacktrace or v1" 2>&1 | tail -8) > gpurun_out/r2v_pytest.txt 2>&1
tail -6 gpurun_out/r2v_pytest.txt | cut -c1-220
timeout 300 python bench.py --schema v1 --steps 10 --warmup 3 --no-cpu --no-host-shim --no-u32 --e2e-steps 2 > gpurun_out/r2v_bench_v1.json 2> gpurun_out/r2v_bench_v1.err
python - <<'PY'
import json
d=json.loads(open("gpurun_out/r2v_bench_v1.json").read().strip().splitlines()[-1])
print(d["ms_per_step"], d["kernel_groups_ms"], d["e2e"]["value"], d.get("v1_stacktrace_record"))
PY
tail -3 gpurun_out/r2v_bench_v1.err
echo done
