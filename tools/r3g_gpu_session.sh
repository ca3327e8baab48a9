for v in direct staged wide; do
PA_HASH_VARIANT=$v timeout 300 python bench.py --config 3 --steps 10 --warmup 3 --no-cpu --no-host-shim --no-u32 --e2e-steps 1 > gpurun_out/r3g_bench_c3_$v.json 2> gpurun_out/r3g_bench_c3_$v.err
done
python - <<'PY'
import json
for v in ("direct","staged","wide"):
    try:
        d=json.loads(open("gpurun_out/r3g_bench_c3_%s.json"%v).read().strip().splitlines()[-1])
        print(v,d["ms_per_step"],d["roofline"]["avg_launch_ms"],d["roofline"]["frac"])
    except Exception as e: print(v,"ERR",e)
PY
echo done
