(time timeout 1200 python -m pytest tests -m gpu -x -q 2>&1 | tail -8) > gpurun_out/r2q_pytest.txt 2>&1
tail -5 gpurun_out/r2q_pytest.txt
timeout 500 python bench.py --steps 10 --warmup 3 --no-host-shim > gpurun_out/r2q_bench.json 2> gpurun_out/r2q_bench.err
PA_NO_EARLY_D2H=1 timeout 300 python bench.py --steps 10 --warmup 3 --no-cpu --no-host-shim > gpurun_out/r2q_bench_noearly.json 2> gpurun_out/r2q_bench_noearly.err
python - <<'PY'
import json
for f in ("r2q_bench","r2q_bench_noearly"):
    try:
        d=json.loads(open("gpurun_out/%s.json"%f).read().strip().splitlines()[-1])
        print(f, d["ms_per_step"], d["e2e"], (d.get("u32_ring") or {}).get("e2e"), d.get("result",{}).get("bit_exact_vs_cpu_port"))
    except Exception as e: print(f,"ERR",e)
PY
tail -3 gpurun_out/r2q_bench.err
echo done
