PA_ONE_GPU=1 timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29715 bench.py --gpus 2 --steps 5 --warmup 3 --merge-rows 4000000 > gpurun_out/r3d_dry_2proc.json 2> gpurun_out/r3d_dry_2proc.err
PA_ONE_GPU=1 timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 3 --master-addr 127.0.0.1 --master-port 29717 bench.py --gpus 3 --steps 3 --warmup 3 --merge-rows 2000000 --no-cpu > gpurun_out/r3d_dry_3proc.json 2> gpurun_out/r3d_dry_3proc.err
python - <<'PY'
import json
for f in ("r3d_dry_2proc","r3d_dry_3proc"):
    try:
        d=json.loads(open("gpurun_out/%s.json"%f).read().strip().splitlines()[-1])
        mb=d.get("mode_b") or {}
        print(f, d["value"], d["e2e"]["value"], mb.get("ms_per_step"), mb.get("value"), (mb.get("e2e") or {}).get("value"), mb.get("error"), mb.get("rows"))
    except Exception as e: print(f,"ERR",e)
PY
tail -c 400 gpurun_out/r3d_dry_3proc.err
echo done
