ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_r1c.csv python bench.py --steps 20 --warmup 3 --no-cpu --e2e-steps 1 > gpurun_out/b_r1c.log 2>&1
N=2 ncu --set full --clock-control none --import-source on -k regex:dhop_col -s 2 -c 2 -o gpurun_out/prof_dhop_r1c -f python scripts/prof_dhop.py > gpurun_out/prof_r1c.log 2>&1
ncu --set full --clock-control none -k regex:stag_dhop -s 3 -c 1 -o gpurun_out/prof_stag_r1c -f python scripts/stag_bench.py 48 6 > gpurun_out/prof_stag.log 2>&1
wc -l gpurun_out/launches_r1c.csv; tail -2 gpurun_out/prof_r1c.log; tail -2 gpurun_out/prof_stag.log
