TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29533"
timeout 420 $TR scripts/mgpu_check.py > gpurun_out/mgpu2.log 2>&1; echo "mgpu rc $?" >> gpurun_out/mgpu2.log
timeout 200 $TR bench.py --gpus 2 --steps 200 --warmup 5 --no-cpu --no-cg --e2e-steps 1 > gpurun_out/bench_n2_col.json 2> gpurun_out/bench_n2_col.err
GB_NO_COL=1 timeout 200 $TR bench.py --gpus 2 --steps 200 --warmup 5 --no-cpu --no-cg --e2e-steps 1 > gpurun_out/bench_n2_nocol.json 2> gpurun_out/bench_n2_nocol.err
grep -E "FAIL|MGPU_CHECK|rc" gpurun_out/mgpu2.log | tail -5; for f in col nocol; do grep -o '"ms_per_step": [0-9.]*' gpurun_out/bench_n2_$f.json | head -1; done
