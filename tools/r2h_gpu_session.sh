(time timeout 300 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "registration_concurrent" 2>&1 | tail -6) > gpurun_out/r2h_pytest.txt 2>&1
timeout 400 python bench.py --config 3 --steps 10 --warmup 3 --no-host-shim > gpurun_out/r2h_bench_config3.json 2> gpurun_out/r2h_bench_config3.err
timeout 400 python bench.py --schema v1 --steps 10 --warmup 3 --no-cpu --no-host-shim > gpurun_out/r2h_bench_v1.json 2> gpurun_out/r2h_bench_v1.err
timeout 300 python bench.py --hash-mode provided --steps 10 --warmup 3 --no-cpu --no-host-shim > gpurun_out/r2h_bench_provided.json 2> gpurun_out/r2h_bench_provided.err
echo done
