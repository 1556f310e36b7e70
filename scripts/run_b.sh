python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "dhop or kernels" 2>&1 | grep -v "Grid : " | tail -15 > gpurun_out/pytest_b.log
tail -4 gpurun_out/pytest_b.log
for n in 8 16 32; do GB_COL_N=$n python bench.py --steps 100 --warmup 5 --no-cpu --no-cg --e2e-steps 1 2>/dev/null | python -c "
import sys,json; d=json.loads(sys.stdin.read()); print('col N=$n', d['ms_per_step'], d['roofline']['frac'], d['clocks'])"; done
