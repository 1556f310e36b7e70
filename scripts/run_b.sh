python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "dhop or kernels" 2>&1 | grep -v "Grid : " | tail -3
python bench.py --steps 100 --warmup 5 --no-cpu --no-cg --e2e-steps 1 2>/dev/null | python -c "
import sys,json; d=json.loads(sys.stdin.read()); print('regs TPL0', d['ms_per_step'], d['roofline']['frac'], d['clocks']['sm_mhz'])"
GB_COL_TPL=1 python bench.py --steps 100 --warmup 5 --no-cpu --no-cg --e2e-steps 1 2>/dev/null | python -c "
import sys,json; d=json.loads(sys.stdin.read()); print('regs TPL1', d['ms_per_step'], d['roofline']['frac'], d['clocks']['sm_mhz'])"
