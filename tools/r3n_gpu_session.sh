(time timeout 1500 python -m pytest tests -m gpu -x -q 2>&1 | tail -6) > gpurun_out/r3n_pytest.txt 2>&1
tail -4 gpurun_out/r3n_pytest.txt
python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" 2>&1 | tail -1
timeout 300 python tests/stress_gpu.py --cases 150 --seed 31 2>&1 | tail -1
echo done
