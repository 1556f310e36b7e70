python -m pytest tests -m gpu -q 2>&1 | grep -v "Grid : " | tail -12 > gpurun_out/pytest_full.log
tail -5 gpurun_out/pytest_full.log
python bench.py > gpurun_out/bench_r1c.json 2> gpurun_out/bench_r1c.err; tail -c 1500 gpurun_out/bench_r1c.json
