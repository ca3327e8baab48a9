PA_COLLECT_PROFILE=1 timeout 300 python bench.py --steps 10 --warmup 3 --no-cpu --no-host-shim --no-u32 --e2e-steps 2 > gpurun_out/r2s_bench.json 2> gpurun_out/r2s_bench.err
grep collect gpurun_out/r2s_bench.err | tail -2
PA_PDL=1 timeout 300 python bench.py --steps 10 --warmup 3 --no-cpu --no-host-shim --e2e-steps 2 > gpurun_out/r2s_bench_pdl.json 2> gpurun_out/r2s_bench_pdl.err
PA_PDL=1 PA_SERIAL=1 timeout 300 python bench.py --steps 10 --warmup 3 --no-cpu --no-host-shim --no-u32 --e2e-steps 2 > gpurun_out/r2s_bench_pdl_serial.json 2> gpurun_out/r2s_bench_pdl_serial.err
PA_SERIAL=1 timeout 300 python bench.py --steps 10 --warmup 3 --no-cpu --no-host-shim --no-u32 --e2e-steps 2 > gpurun_out/r2s_bench_serial.json 2> gpurun_out/r2s_bench_serial.err
(PA_PDL=1 timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_merge.py -m gpu -x -q -k "not full and not early_copy and not multi_process" 2>&1 | tail -4) > gpurun_out/r2s_pytest_pdl.txt 2>&1
tail -3 gpurun_out/r2s_pytest_pdl.txt
python - <<'PY'
import json
for f in ("r2s_bench","r2s_bench_pdl","r2s_bench_pdl_serial","r2s_bench_serial"):
    try:
        d=json.loads(open("gpurun_out/%s.json"%f).read().strip().splitlines()[-1])
        print(f, d["ms_per_step"], d["kernel_groups_ms"], d["e2e"]["value"])
    except Exception as e: print(f,"ERR",e)
PY
echo done
