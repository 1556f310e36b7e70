"""Improved staggered Dhop fp32 on N B200 (BASELINE configs[4]: 48^4 with the Naik 3-hop stencil on 1/8 B200), one rank per GPU:
   python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29544 scripts/stag_bench_mgpu.py [L] [ncall] [weak]
Default = strong scaling (global L^4 split as 1.1.2.4 on 8 ranks, 1.1.2.2 on 4, 1.1.1.2 on 2); "weak" keeps L^4 per rank.
Time = CUDA events on the library's stream, max over ranks; GFlop/s with 1146 flop/site (ref: benchmarks/Benchmark_staggered.cc:105)."""
import json, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import torch.distributed as dist
import grid_b200 as gb

rank, world, lrank = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1)), int(os.environ.get("LOCAL_RANK", 0))
torch.cuda.set_device(lrank)
if world > 1:
    dist.init_process_group("nccl", device_id=torch.device("cuda", lrank))
L = int(sys.argv[1]) if len(sys.argv) > 1 else 48
ncall = int(sys.argv[2]) if len(sys.argv) > 2 else 200
weak = len(sys.argv) > 3 and sys.argv[3] == "weak"
mpi = {1: (1, 1, 1, 1), 2: (1, 1, 1, 2), 4: (1, 1, 2, 2), 8: (1, 1, 2, 4)}[world]
gdims = tuple(L * m for m in mpi) if weak else (L,) * 4
ctx = gb.Context(lrank)
uid = [gb.Context.unique_id() if rank == 0 else None]
if world > 1:
    dist.broadcast_object_list(uid, src=0)
ctx.comm_init(rank, world, uid[0])
grid = gb.GridCartesian(ctx, gdims, mpi)
for prec, name, w in ((gb.F32, "fp32", 4), (gb.F64, "fp64", 8)):
    U = gb.LatticeGaugeField(grid, prec).random(1)
    D = gb.ImprovedStaggeredFermion(U, U, grid, 0.1)
    del U
    src, out = gb.LatticeStaggeredFermion(grid, 1, prec).random(2), gb.LatticeStaggeredFermion(grid, 1, prec)
    for _ in range(5):
        D.Dhop(src, out, 0)
    ctx.synchronize()
    if world > 1:
        dist.barrier()
    ctx.timer_start()
    for _ in range(ncall):
        D.Dhop(src, out, 0)
    ms = ctx.timer_stop() / ncall
    if world > 1:
        t = torch.tensor([ms], dtype=torch.float64, device=f"cuda:{lrank}")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms = t.item()
    sites = 1
    for g in gdims:
        sites *= g
    if rank == 0:
        print(json.dumps({"op": f"ImprovedStaggeredFermion::Dhop {name}", "gdims": gdims, "mpi": mpi, "n_gpus": world, "scaling": "weak" if weak else "strong",
                          "ms": ms, "gflops_total": 1146.0 * sites / ms / 1e6, "alg_GBs_per_gpu": (2 * 6 * w + 16 * 18 * w) * sites / world / ms / 1e6}), flush=True)
    del D, src, out
if world > 1:
    dist.destroy_process_group()
