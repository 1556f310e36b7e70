GB_COL_NT=2 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "dhop or kernels" 2>&1 | grep -v "Grid : " | tail -3
GB_COL_NT=2 python bench.py --steps 100 --warmup 5 --no-cpu --no-cg --e2e-steps 1 2>/dev/null | python -c "
import sys,json; d=json.loads(sys.stdin.read()); print('NT2', d['ms_per_step'], d['roofline']['frac'], d['clocks']['sm_mhz'])"
python bench.py --steps 100 --warmup 5 --no-cpu --e2e-steps 1 2>/dev/null | python -c "
import sys,json; d=json.loads(sys.stdin.read()); print('NT1', d['ms_per_step'], d['roofline']['frac'], d['clocks']['sm_mhz'], 'CG', d['cg']['time_to_solution_s'], d['cg']['inner_iterations'])"
