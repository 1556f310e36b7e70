N=${1:-4}
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29533"
timeout 500 $TR scripts/mgpu_check.py > gpurun_out/mgpu$N.log 2>&1; echo "mgpu rc $?" >> gpurun_out/mgpu$N.log
timeout 300 $TR bench.py --gpus $N --steps 200 --warmup 5 --no-cpu --e2e-steps 1 > gpurun_out/bench_n$N.json 2> gpurun_out/bench_n$N.err
grep -E "FAIL|MGPU_CHECK|rc" gpurun_out/mgpu$N.log | tail -5; grep -o '"ms_per_step": [0-9.]*' gpurun_out/bench_n$N.json | head -1; grep -o '"cg": {[^}]*}' gpurun_out/bench_n$N.json
