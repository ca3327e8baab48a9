for cs in 1 2; do
PA_COPY_STREAMS=$cs timeout 300 python bench.py --steps 3 --warmup 3 --no-cpu --no-host-shim --e2e-steps 4 > gpurun_out/r3j_bench_cs$cs.json 2> gpurun_out/r3j_bench_cs$cs.err
done
python - <<'PY'
import json
for f in ("cs1","cs2"):
    try:
        d=json.loads(open("gpurun_out/r3j_bench_%s.json"%f).read().strip().splitlines()[-1])
        print(f, d["e2e"]["value"], d["e2e"]["stages_ms_last_step"], d["u32_ring"]["e2e"]["value"])
    except Exception as e: print(f,"ERR",e, open("gpurun_out/r3j_bench_%s.err"%f).read()[-300:])
PY
echo done
