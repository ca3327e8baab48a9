(time timeout 1500 python -m pytest tests -m gpu -x -q 2>&1 | tail -6) > gpurun_out/r3h_pytest.txt 2>&1
tail -4 gpurun_out/r3h_pytest.txt
python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" > gpurun_out/r3h_smoke.txt 2>&1; tail -1 gpurun_out/r3h_smoke.txt
(time timeout 600 python bench.py --steps 20 --warmup 5) > gpurun_out/r3h_bench_default.json 2> gpurun_out/r3h_bench_default.err
PA_EARLY_D2H_ALWAYS=1 timeout 600 python tests/stress_gpu.py --cases 300 --seed 21 > gpurun_out/r3h_stress.txt 2>&1; tail -1 gpurun_out/r3h_stress.txt
python - <<'PY'
import json
d=json.loads(open("gpurun_out/r3h_bench_default.json").read().strip().splitlines()[-1])
print(d["steps"], d["ms_per_step"], d["e2e"]["value"], d["u32_ring"]["e2e"]["value"], d["result"]["bit_exact_vs_cpu_port"], d["cpu_baseline"]["value"], d["roofline"]["frac"])
print(d["host_shim"]["e2e_submit"]["value"], d["host_shim"]["report_trace_event_handle"]["ingest_samples_per_s"])
PY
echo done
