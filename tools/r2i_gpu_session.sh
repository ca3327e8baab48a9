(time timeout 900 python -m pytest tests/test_merge.py -m gpu -x -q -k "multi_process" 2>&1 | tail -8) > gpurun_out/r2i_pytest_merge_mp.txt 2>&1
PA_ONE_GPU=1 timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29611 bench.py --gpus 2 --steps 5 --warmup 3 --merge-rows 4000000 > gpurun_out/r2i_dry_shm.json 2> gpurun_out/r2i_dry_shm.err
PA_ONE_GPU=1 timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29631 bench.py --gpus 2 --steps 5 --warmup 3 --merge-rows 4000000 --merge-transport host > gpurun_out/r2i_dry_gloo.json 2> gpurun_out/r2i_dry_gloo.err
tail -c 600 gpurun_out/r2i_dry_shm.err
echo done
