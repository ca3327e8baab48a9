(time timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "early_copy_out or stacktrace or widepf3" 2>&1 | tail -6) > gpurun_out/r3a_pytest.txt 2>&1
tail -5 gpurun_out/r3a_pytest.txt
for v in widepf3 wide; do
PA_HASH_VARIANT=$v timeout 300 python bench.py --steps 10 --warmup 3 --no-cpu --no-host-shim --no-u32 --e2e-steps 1 > gpurun_out/r3a_bench_$v.json 2> gpurun_out/r3a_bench_$v.err
PA_HASH_VARIANT=$v timeout 300 python bench.py --config 3 --steps 10 --warmup 3 --no-cpu --no-host-shim --no-u32 --e2e-steps 1 > gpurun_out/r3a_bench_c3_$v.json 2> gpurun_out/r3a_bench_c3_$v.err
done
python - <<'PY'
import json
for v in ("widepf3","wide"):
    for c in ("","c3_"):
        try:
            d=json.loads(open("gpurun_out/r3a_bench_%s%s.json"%(c,v)).read().strip().splitlines()[-1])
            print(v,c,d["ms_per_step"],d["roofline"]["avg_launch_ms"],d["roofline"]["frac"])
        except Exception as e: print(v,c,"ERR",e)
PY
echo done
