TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29533"
MGPU_MPI_INDEX=0 timeout 300 $TR scripts/mgpu_check.py > gpurun_out/mgpu2_sf.log 2>&1; echo "mgpu rc $?" >> gpurun_out/mgpu2_sf.log
grep -E "FAIL|MGPU_CHECK|rc" gpurun_out/mgpu2_sf.log | tail -4
timeout 200 $TR bench.py --gpus 2 --steps 200 --warmup 5 --no-cpu --e2e-steps 1 > gpurun_out/bench_n2_sf.json 2> gpurun_out/bench_n2_sf.err
echo semifused Dhop $(grep -o '"ms_per_step": [0-9.]*' gpurun_out/bench_n2_sf.json | head -1) CG $(grep -o '"time_to_solution_s": [0-9.]*' gpurun_out/bench_n2_sf.json)
