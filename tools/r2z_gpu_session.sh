PA_EARLY_D2H_ALWAYS=1 timeout 1500 compute-sanitizer --tool memcheck --error-exitcode 9 python tools/sanitize_small.py > gpurun_out/r2z_memcheck.txt 2>&1
echo "memcheck rc=$?" >> gpurun_out/r2z_memcheck.txt
tail -6 gpurun_out/r2z_memcheck.txt
PA_EARLY_D2H_ALWAYS=1 timeout 1500 compute-sanitizer --tool racecheck --error-exitcode 9 python tools/sanitize_small.py > gpurun_out/r2z_racecheck.txt 2>&1
echo "racecheck rc=$?" >> gpurun_out/r2z_racecheck.txt
tail -6 gpurun_out/r2z_racecheck.txt
echo done
