(time timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "hash_kernel_variants or narrow or config2_full or config3_full or ragged or large_batch" 2>&1 | tail -4) > gpurun_out/r3m_pytest.txt 2>&1
tail -3 gpurun_out/r3m_pytest.txt
(time timeout 600 python bench.py --steps 20 --warmup 5) > gpurun_out/r3m_bench_default.json 2> gpurun_out/r3m_bench_default.err
python - <<'PY'
import json
d=json.loads(open("gpurun_out/r3m_bench_default.json").read().strip().splitlines()[-1])
print(d["steps"], d["ms_per_step"], d["roofline"]["avg_launch_ms"], d["roofline"]["frac"], d["e2e"]["value"], d["u32_ring"]["e2e"]["value"], d["u32_ring"]["ms_per_step"], d["u32_ring"]["hash_kernel"]["ms"], d["result"]["bit_exact_vs_cpu_port"])
PY
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r3m_launches.csv python bench.py --steps 2 --warmup 1 --no-cpu --no-host-shim --no-u32 --e2e-steps 1 > gpurun_out/r3m_ncu_bench.log 2>&1
timeout 500 ncu --set full --clock-control none --import-source on -k regex:"k_hash_insert_wide|k_ree_col|k_header" -s 6 -c 4 -o gpurun_out/r3m_top python bench.py --steps 2 --warmup 1 --no-cpu --no-host-shim --no-u32 --e2e-steps 1 > gpurun_out/r3m_ncu_full.log 2>&1
echo done
