(time timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "stacktrace or v1 or mode_b_merged" 2>&1 | tail -15) > gpurun_out/r2t_pytest.txt 2>&1
tail -15 gpurun_out/r2t_pytest.txt
PA_COLLECT_PROFILE=1 timeout 300 python bench.py --steps 10 --warmup 3 --no-cpu --no-host-shim --no-u32 --e2e-steps 2 > gpurun_out/r2t_bench.json 2> gpurun_out/r2t_bench.err
grep collect gpurun_out/r2t_bench.err | tail -2
PA_FORK_EARLY=1 timeout 300 python bench.py --steps 10 --warmup 3 --no-cpu --no-host-shim --no-u32 --e2e-steps 1 > gpurun_out/r2t_bench_forkearly.json 2> gpurun_out/r2t_bench_forkearly.err
timeout 300 python bench.py --schema v1 --steps 10 --warmup 3 --no-cpu --no-host-shim --no-u32 --e2e-steps 2 > gpurun_out/r2t_bench_v1.json 2> gpurun_out/r2t_bench_v1.err
python - <<'PY'
import json
for f in ("r2t_bench","r2t_bench_forkearly","r2t_bench_v1"):
    try:
        d=json.loads(open("gpurun_out/%s.json"%f).read().strip().splitlines()[-1])
        print(f, d["ms_per_step"], d["kernel_groups_ms"], d["e2e"]["value"], d.get("v1_stacktrace_record"))
    except Exception as e: print(f,"ERR",e)
PY
tail -3 gpurun_out/r2t_bench_v1.err
echo done
