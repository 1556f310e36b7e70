"""Halo microbenchmark (the reference's benchmarks/Benchmark_comms.cc:105-162 shape): the face exchange of one fp32
DWF Dhop at 32^4 x Ls16 per GPU on its own -- fused project+send of L^3 x Ls half-spinor packets in every split direction, both
senses concurrently, then arrival -- timed with CUDA events over N calls, max over ranks.
   python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port 29555 scripts/halo_bench.py [L] [ncall]
Prints one JSON line per (mpi, path): bytes sent per rank per call, ms, GB/s out of each GPU (= in, by symmetry), and the same
number per direction (/ 2 per split dimension)."""
import json, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import torch.distributed as dist
import grid_b200 as gb

rank, world, lrank = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(lrank)
dist.init_process_group("nccl", device_id=torch.device("cuda", lrank))
L = int(sys.argv[1]) if len(sys.argv) > 1 else 32
ncall = int(sys.argv[2]) if len(sys.argv) > 2 else 200
Ls = 16
mpi = {2: (1, 1, 1, 2), 4: (1, 1, 2, 2), 8: (1, 1, 2, 4)}[world]
ctx = gb.Context(lrank)
uid = [gb.Context.unique_id() if rank == 0 else None]
dist.broadcast_object_list(uid, src=0)
ctx.comm_init(rank, world, uid[0])
grid = gb.GridCartesian(ctx, tuple(L * m for m in mpi), mpi)
U = gb.LatticeGaugeField(grid, gb.F32).random(1)
D = gb.DomainWallFermion(U, grid, Ls, 0.1, 1.8)
src, out = gb.LatticeFermion(grid, Ls, gb.F32).random(2), gb.LatticeFermion(grid, Ls, gb.F32)
path = "nccl send/recv" if os.environ.get("GB_NO_P2P") else "peer-to-peer stores"
for _ in range(5):
    nbytes = D.halo_exchange(src)
ctx.synchronize(); dist.barrier()
ctx.timer_start()
for _ in range(ncall):
    D.halo_exchange(src)
ms = ctx.timer_stop() / ncall
t = torch.tensor([ms], dtype=torch.float64, device=f"cuda:{lrank}")
dist.all_reduce(t, op=dist.ReduceOp.MAX)
ms = t.item()
# the same hop with its exchange, for the share the exchange has of it
for _ in range(5):
    D.Dhop(src, out, 0)
ctx.synchronize(); dist.barrier()
ctx.timer_start()
for _ in range(ncall):
    D.Dhop(src, out, 0)
hop = ctx.timer_stop() / ncall
t = torch.tensor([hop], dtype=torch.float64, device=f"cuda:{lrank}")
dist.all_reduce(t, op=dist.ReduceOp.MAX)
if rank == 0:
    nsplit = sum(1 for m in mpi if m > 1)
    print(json.dumps({"bench": "halo exchange of one fp32 DWF Dhop", "local": [L] * 4, "Ls": Ls, "mpi": mpi, "n_gpus": world, "path": path,
                      "bytes_sent_per_rank": nbytes, "ms": ms, "GBs_out_per_gpu": nbytes / ms / 1e6,
                      "GBs_per_direction": nbytes / ms / 1e6 / (2 * nsplit), "dhop_ms_with_exchange": t.item()}), flush=True)
dist.destroy_process_group()
