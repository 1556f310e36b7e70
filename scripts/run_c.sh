python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | grep -v "Grid : " | tail -2
python -m pytest tests -m gpu -q -x 2>&1 | grep -v "Grid : " | tail -6 > gpurun_out/pytest_full.log
tail -3 gpurun_out/pytest_full.log
python bench.py > gpurun_out/bench_r1d.json 2> gpurun_out/bench_r1d.err; python -c "
import json; d=json.load(open('gpurun_out/bench_r1d.json')); print(d['ms_per_step'], d['roofline']['frac'], d['e2e']['value'], d['cg']['time_to_solution_s'], d['cpu_baseline']['value'], d['cpu_baseline']['kind'], d['clocks'])"
