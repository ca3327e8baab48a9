(time timeout 1500 python -m pytest tests -m gpu -x -q 2>&1 | tail -6) > gpurun_out/r3c_pytest.txt 2>&1
tail -4 gpurun_out/r3c_pytest.txt
python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" > gpurun_out/r3c_smoke.txt 2>&1; tail -1 gpurun_out/r3c_smoke.txt
(time timeout 600 python bench.py) > gpurun_out/r3c_bench_default.json 2> gpurun_out/r3c_bench_default.err
timeout 400 python bench.py --config 3 --steps 10 --warmup 3 --no-host-shim > gpurun_out/r3c_bench_config3.json 2> gpurun_out/r3c_bench_config3.err
timeout 600 python bench.py --stream 4 > gpurun_out/r3c_stream.json 2> gpurun_out/r3c_stream.err
PA_ONE_GPU=1 timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29713 bench.py --gpus 2 --steps 5 --warmup 3 --merge-rows 4000000 > gpurun_out/r3c_dry_2proc.json 2> gpurun_out/r3c_dry_2proc.err
python - <<'PY'
import json
def last(f): return json.loads(open("gpurun_out/%s.json"%f).read().strip().splitlines()[-1])
for f in ("r3c_bench_default","r3c_bench_config3","r3c_stream","r3c_dry_2proc"):
    try:
        d=last(f)
        print(f, d.get("value"), d.get("ms_per_step"), (d.get("e2e") or {}).get("value"), ((d.get("u32_ring") or {}).get("e2e") or {}).get("value"), (d.get("roofline") or {}).get("frac"), (d.get("result") or {}).get("bit_exact_vs_cpu_port"), (d.get("mode_b") or {}).get("ms_per_step"), (d.get("mode_b") or {}).get("error"))
    except Exception as e: print(f,"ERR",e)
PY
echo done
