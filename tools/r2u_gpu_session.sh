(time timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "stacktrace or v1 or mode_b_merged" 2>&1 | tail -30) > gpurun_out/r2u_pytest.txt 2>&1
tail -30 gpurun_out/r2u_pytest.txt | cut -c1-220
echo done
