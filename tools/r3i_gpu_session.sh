(time timeout 600 python bench.py --steps 5 --warmup 3 --no-cpu) > gpurun_out/r3i_bench.json 2> gpurun_out/r3i_bench.err
python - <<'PY'
import json
d=json.loads(open("gpurun_out/r3i_bench.json").read().strip().splitlines()[-1])
print(d["ms_per_step"], d["e2e"]["value"])
print(json.dumps({k:v for k,v in d["host_shim"].items() if k.startswith("e2e_submit") or k=="error"})[:900])
PY
tail -c 300 gpurun_out/r3i_bench.err
echo done
