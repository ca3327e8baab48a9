(time timeout 600 python -m pytest tests/test_gpu_parity.py tests/test_merge.py -m gpu -x -q -k "narrow or single_ring or label_names or merge" 2>&1 | tail -8) > gpurun_out/r2f_pytest.txt 2>&1
timeout 400 python bench.py --steps 10 --warmup 3 --no-cpu > gpurun_out/r2f_bench.json 2> gpurun_out/r2f_bench.err
PA_REE_BLOCKS_PER_SM=6 timeout 200 python bench.py --steps 10 --warmup 3 --no-cpu --no-host-shim --no-u32 > gpurun_out/r2f_bench_ree6.json 2> gpurun_out/r2f_bench_ree6.err
PA_REE_BLOCKS_PER_SM=4 timeout 200 python bench.py --steps 10 --warmup 3 --no-cpu --no-host-shim --no-u32 > gpurun_out/r2f_bench_ree4.json 2> gpurun_out/r2f_bench_ree4.err
timeout 600 python bench.py --stream 4 > gpurun_out/r2f_stream1.json 2> gpurun_out/r2f_stream1.err
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r2f_launches.csv python bench.py --steps 2 --warmup 1 --no-cpu --no-host-shim --no-u32 --e2e-steps 1 > gpurun_out/r2f_ncu_bench.log 2>&1
timeout 500 ncu --set full --clock-control none --import-source on -k regex:"k_hash_insert_wide|k_ree_col|k_header" -s 6 -c 4 -o gpurun_out/r2f_top python bench.py --steps 2 --warmup 1 --no-cpu --no-host-shim --no-u32 --e2e-steps 1 > gpurun_out/r2f_ncu_full.log 2>&1
PA_HASH_VARIANT=bulk6x2 timeout 400 ncu --set full --clock-control none --import-source on -k regex:"k_hash_insert_bulk" -s 1 -c 1 -o gpurun_out/r2f_bulk python bench.py --steps 2 --warmup 1 --no-cpu --no-host-shim --no-u32 --e2e-steps 1 > gpurun_out/r2f_ncu_bulk.log 2>&1
echo done
