PA_EARLY_D2H_ALWAYS=1 timeout 900 python tests/stress_gpu.py --cases 400 --seed 7 > gpurun_out/r2y_stress.txt 2>&1
tail -2 gpurun_out/r2y_stress.txt
timeout 600 python tests/stress_gpu.py --cases 300 --seed 11 > gpurun_out/r2y_stress2.txt 2>&1
tail -2 gpurun_out/r2y_stress2.txt
echo done
