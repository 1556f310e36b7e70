TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29533"
MGPU_MPI_INDEX=0 timeout 400 $TR scripts/mgpu_check.py > gpurun_out/mgpu8.log 2>&1; echo "mgpu rc $?" >> gpurun_out/mgpu8.log
timeout 200 $TR bench.py --gpus 8 --steps 100 --warmup 5 --no-cpu --no-cg --e2e-steps 1 > gpurun_out/bench_n8.json 2> gpurun_out/bench_n8.err
grep -E "FAIL|MGPU_CHECK|rc" gpurun_out/mgpu8.log | tail -5; grep -o '"ms_per_step": [0-9.]*' gpurun_out/bench_n8.json | head -1; tail -3 gpurun_out/bench_n8.err
