timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r2x_launches.csv python bench.py --steps 2 --warmup 1 --no-cpu --no-host-shim --no-u32 --e2e-steps 1 > gpurun_out/r2x_ncu_bench.log 2>&1
timeout 500 ncu --set full --clock-control none --import-source on -k regex:"k_hash_insert_wide|k_ree_col|k_header" -s 6 -c 4 -o gpurun_out/r2x_top python bench.py --steps 2 --warmup 1 --no-cpu --no-host-shim --no-u32 --e2e-steps 1 > gpurun_out/r2x_ncu_full.log 2>&1
tail -2 gpurun_out/r2x_ncu_full.log
(time timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "early_copy_out or is_an_lru" --durations=3 2>&1 | tail -8) > gpurun_out/r2x_pytest.txt 2>&1
tail -8 gpurun_out/r2x_pytest.txt
echo done
