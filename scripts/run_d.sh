python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "dhop_host" 2>&1 | grep -v "Grid : " | tail -8
python bench.py --steps 50 --warmup 5 --no-cpu --no-cg --e2e-steps 5 2>/dev/null | python -c "
import sys,json; d=json.loads(sys.stdin.read()); print(d['ms_per_step'], d['e2e'])"
