# final validation of the round: whole GPU suite, default bench (all legs), reference arm, config 3 / v1 / provided records, streaming, the N>1 flow on one GPU
(time timeout 1500 python -m pytest tests -m gpu -x -q 2>&1 | tail -8) > gpurun_out/r2w_pytest.txt 2>&1
tail -5 gpurun_out/r2w_pytest.txt
(time timeout 600 python bench.py) > gpurun_out/r2w_bench_default.json 2> gpurun_out/r2w_bench_default.err
tail -c 300 gpurun_out/r2w_bench_default.err
(time timeout 600 python bench.py --impl reference --steps 1 --warmup 1) > gpurun_out/r2w_bench_reference.json 2> gpurun_out/r2w_bench_reference.err
timeout 400 python bench.py --config 3 --steps 10 --warmup 3 --no-host-shim > gpurun_out/r2w_bench_config3.json 2> gpurun_out/r2w_bench_config3.err
timeout 400 python bench.py --schema v1 --steps 10 --warmup 3 --no-cpu --no-host-shim > gpurun_out/r2w_bench_v1.json 2> gpurun_out/r2w_bench_v1.err
timeout 300 python bench.py --hash-mode provided --steps 10 --warmup 3 --no-cpu --no-host-shim > gpurun_out/r2w_bench_provided.json 2> gpurun_out/r2w_bench_provided.err
timeout 600 python bench.py --stream 4 > gpurun_out/r2w_stream.json 2> gpurun_out/r2w_stream.err
PA_ONE_GPU=1 timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29711 bench.py --gpus 2 --steps 5 --warmup 3 --merge-rows 4000000 > gpurun_out/r2w_dry_2proc.json 2> gpurun_out/r2w_dry_2proc.err
python - <<'PY'
import json
def last(f):
    return json.loads(open("gpurun_out/%s.json"%f).read().strip().splitlines()[-1])
for f in ("r2w_bench_default","r2w_bench_reference","r2w_bench_config3","r2w_bench_v1","r2w_bench_provided","r2w_stream","r2w_dry_2proc"):
    try:
        d=last(f)
        print(f, d.get("value"), d.get("ms_per_step"), (d.get("e2e") or {}).get("value"), (d.get("roofline") or {}).get("frac"), (d.get("result") or {}).get("bit_exact_vs_cpu_port"), (d.get("mode_b") or {}).get("ms_per_step"), (d.get("mode_b") or {}).get("error"))
    except Exception as e: print(f,"ERR",e)
PY
echo done
