#!/bin/bash
# GPU call G: the whole bench line (headline + cg + e2e_cg + config4 + config5 + cpu leg) on one GPU, then the reference arm.
set -u
out=gpurun_out/r2g; mkdir -p $out
nproc > $out/host.txt; free -g >> $out/host.txt; nvidia-smi topo -m >> $out/host.txt 2>&1
( time timeout 900 python bench.py ) > $out/bench_n1.json 2> $out/bench_n1.err
echo "bench rc $?"; tail -c 3000 $out/bench_n1.json; tail -5 $out/bench_n1.err
( time timeout 600 python bench.py --impl reference --steps 3 --warmup 1 ) > $out/bench_ref.json 2> $out/bench_ref.err
echo "ref rc $?"; tail -c 1500 $out/bench_ref.json
