df -h /dev/shm | tail -1
(time timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "hash_kernel_variants or narrow_frame" 2>&1 | tail -6) > gpurun_out/r2k_pytest.txt 2>&1
cat gpurun_out/r2k_pytest.txt | tail -4
for v in wide widepf; do
PA_HASH_VARIANT=$v timeout 400 python bench.py --steps 10 --warmup 3 --no-cpu --no-host-shim > gpurun_out/r2k_bench_$v.json 2> gpurun_out/r2k_bench_$v.err
PA_HASH_VARIANT=$v timeout 400 python bench.py --config 3 --steps 10 --warmup 3 --no-cpu --no-host-shim > gpurun_out/r2k_bench_c3_$v.json 2> gpurun_out/r2k_bench_c3_$v.err
done
python - <<'PY'
import json
for v in ("wide","widepf"):
    for c in ("","c3_"):
        try:
            d=json.loads(open("gpurun_out/r2k_bench_%s%s.json"%(c,v)).read().strip().splitlines()[-1])
            print(v,c,d["ms_per_step"],d["roofline"]["avg_launch_ms"],d["roofline"]["frac"],d["e2e"]["value"], (d.get("u32_ring") or {}).get("ms_per_step"), (d.get("u32_ring") or {}).get("hash_ms"))
        except Exception as e: print(v,c,"ERR",e)
PY
echo done
