(time timeout 1500 python -m pytest tests -m gpu -x -q 2>&1 | tail -6) > gpurun_out/r3k_pytest.txt 2>&1
tail -4 gpurun_out/r3k_pytest.txt
python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" 2>&1 | tail -1
(time timeout 600 python bench.py) > gpurun_out/r3k_bench_default.json 2> gpurun_out/r3k_bench_default.err
python - <<'PY'
import json
d=json.loads(open("gpurun_out/r3k_bench_default.json").read().strip().splitlines()[-1])
print(d["steps"], d["ms_per_step"], d["e2e"]["value"], d["u32_ring"]["e2e"]["value"], d["result"]["bit_exact_vs_cpu_port"], d["cpu_baseline"]["value"], d["roofline"]["frac"], d["gpu_launches"], d["clocks"])
PY
echo done
