TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29533"
timeout 420 $TR scripts/mgpu_check.py > gpurun_out/mgpu2_fused.log 2>&1; echo "mgpu rc $?" >> gpurun_out/mgpu2_fused.log
GB_PROFILE=1 timeout 200 $TR bench.py --gpus 2 --steps 200 --warmup 5 --no-cpu --no-cg --e2e-steps 1 > gpurun_out/bench_n2_fused.json 2> gpurun_out/bench_n2_fused.err
timeout 200 $TR bench.py --gpus 2 --steps 200 --warmup 5 --no-cpu --no-cg --e2e-steps 1 > gpurun_out/bench_n2_fused_b.json 2>> gpurun_out/bench_n2_fused.err
GB_NO_FUSED=1 timeout 200 $TR bench.py --gpus 2 --steps 200 --warmup 5 --no-cpu --no-cg --e2e-steps 1 > gpurun_out/bench_n2_nofused.json 2> gpurun_out/bench_n2_nofused.err
tail -3 gpurun_out/mgpu2_fused.log; cat gpurun_out/bench_n2_fused_b.json gpurun_out/bench_n2_nofused.json | cut -c1-300
