TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29533"
for v in split nosplit; do
  if [ $v = nosplit ]; then export GB_NO_SPLIT=1; else unset GB_NO_SPLIT; fi
  timeout 200 $TR bench.py --gpus 2 --steps 200 --warmup 5 --no-cpu --e2e-steps 1 > gpurun_out/bench_n2_$v.json 2> gpurun_out/bench_n2_$v.err
  timeout 200 $TR bench.py --gpus 2 --steps 200 --warmup 5 --no-cpu --no-cg --e2e-steps 1 --op DhopEO > gpurun_out/bench_n2_eo_$v.json 2>> gpurun_out/bench_n2_$v.err
  echo $v Dhop $(grep -o '"ms_per_step": [0-9.]*' gpurun_out/bench_n2_$v.json | head -1) DhopEO $(grep -o '"ms_per_step": [0-9.]*' gpurun_out/bench_n2_eo_$v.json | head -1) CG $(grep -o '"time_to_solution_s": [0-9.]*' gpurun_out/bench_n2_$v.json)
done
