python -m pytest tests/test_golden.py tests/test_gpu_vs_reference.py tests/test_gpu_staggered.py -m gpu -q 2>&1 | grep -v "Grid : " | tail -150 > gpurun_out/pytest_a.log
python scripts/stag_bench.py 48 200 > gpurun_out/stag_bench.log 2>&1
tail -5 gpurun_out/pytest_a.log; cat gpurun_out/stag_bench.log
