PA_HASH_VARIANT=tma24x2 timeout 400 ncu --set full --clock-control none --import-source on -k regex:"k_hash_insert_tma" -s 1 -c 1 -o gpurun_out/r2m_tma24x2 python bench.py --steps 2 --warmup 1 --no-cpu --no-host-shim --no-u32 --e2e-steps 1 > gpurun_out/r2m_ncu.log 2>&1
tail -3 gpurun_out/r2m_ncu.log
echo done
