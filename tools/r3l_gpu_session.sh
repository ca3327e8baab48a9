(time timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "hash_kernel_variants and widemp" 2>&1 | tail -4) > gpurun_out/r3l_pytest.txt 2>&1
tail -3 gpurun_out/r3l_pytest.txt
for v in widemp wide widemp wide; do
PA_HASH_VARIANT=$v timeout 300 python bench.py --steps 20 --warmup 5 --no-cpu --no-host-shim --no-u32 --e2e-steps 1 > gpurun_out/r3l_bench_$v.json 2> gpurun_out/r3l_bench_$v.err
python - <<PY
import json
d=json.loads(open("gpurun_out/r3l_bench_$v.json").read().strip().splitlines()[-1])
print("$v", d["ms_per_step"], d["roofline"]["avg_launch_ms"], d["roofline"]["frac"])
PY
done
for v in widemp wide; do
PA_HASH_VARIANT=$v timeout 300 python bench.py --config 3 --steps 20 --warmup 5 --no-cpu --no-host-shim --no-u32 --e2e-steps 1 > gpurun_out/r3l_bench_c3_$v.json 2> gpurun_out/r3l_bench_c3_$v.err
python - <<PY
import json
d=json.loads(open("gpurun_out/r3l_bench_c3_$v.json").read().strip().splitlines()[-1])
print("c3 $v", d["ms_per_step"], d["roofline"]["avg_launch_ms"], d["roofline"]["frac"])
PY
done
echo done
