(time timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "early_copy_out" 2>&1 | tail -6) > gpurun_out/r2r_pytest.txt 2>&1
tail -5 gpurun_out/r2r_pytest.txt
PA_COLLECT_PROFILE=1 timeout 300 python bench.py --steps 3 --warmup 3 --no-cpu --no-host-shim --no-u32 > gpurun_out/r2r_bench.json 2> gpurun_out/r2r_bench.err
grep collect gpurun_out/r2r_bench.err | tail -4
echo done
