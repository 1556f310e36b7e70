"""Benchmark_usqcd-shaped table on the B200-native path (ref: benchmarks/Benchmark_usqcd.cc:391-554 DWF4, :556-707 staggered,
:203-262 memory bandwidth): fp32 DhopEO at local L^4 per rank for L in 8,12,16,24,32 -- Wilson (Ls = 1), DWF4 (Ls = 12),
improved staggered -- in the reference's flop conventions (1344 flop per 5D site / 2 for the checkerboard hop, :517-519; 1146
per site / 2 for staggered, :674), plus a stream triad.  These are the rows BASELINE.md quotes for Booster (4 x A100 per node:
DWF4 11487, Wilson 5726, Staggered 2518 GFlop/s per NODE at L = 32) and Frontier.
   python scripts/benchmark_usqcd.py                       (one GPU)
   python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port 29577 scripts/benchmark_usqcd.py
Prints one JSON line per (operator, L): GFlop/s per GPU, max over ranks."""
import json, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import grid_b200 as gb

rank, world, lrank = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1)), int(os.environ.get("LOCAL_RANK", 0))
dist = None
if world > 1:
    import torch
    import torch.distributed as dist
    torch.cuda.set_device(lrank)
    dist.init_process_group("nccl", device_id=torch.device("cuda", lrank))
mpi = {1: (1, 1, 1, 1), 2: (1, 1, 1, 2), 4: (1, 1, 2, 2), 8: (1, 1, 2, 4)}[world]
ctx = gb.Context(lrank)
uid = [gb.Context.unique_id() if rank == 0 else None]
if dist:
    dist.broadcast_object_list(uid, src=0)
ctx.comm_init(rank, world, uid[0])
L_list = [int(x) for x in sys.argv[1:]] or [8, 12, 16, 24, 32]
FPS = 3 * (6 + 2 * 8) * 4 * 4 + 2 * 4 * 3 * 4 + 2 * 4 * 3 * 4 * 2      # 1344, ref :517
BOOSTER_PER_GPU = {"DWF4": {16: 8464 / 4, 24: 10139 / 4, 32: 11487 / 4}, "Wilson": {32: 5726 / 4}, "Staggered": {32: 2518 / 4}}


def timed(fn, ncall):
    for _ in range(5):
        fn()
    ctx.synchronize()
    if dist:
        dist.barrier()
    ctx.timer_start()
    for _ in range(ncall):
        fn()
    ms = ctx.timer_stop() / ncall
    if dist:
        import torch
        t = torch.tensor([ms], dtype=torch.float64, device=f"cuda:{lrank}")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms = t.item()
    return ms


for L in L_list:
    grid = gb.GridCartesian(ctx, tuple(L * m for m in mpi), mpi)
    U = gb.LatticeGaugeField(grid, gb.F32).random(1)
    V4 = L ** 4
    ncall = 400 if L <= 16 else 200
    cases = [("Wilson", lambda: gb.WilsonFermion(U, grid, 0.1), 1, FPS), ("DWF4", lambda: gb.DomainWallFermion(U, grid, 12, 0.1, 1.8), 12, FPS)]
    if L >= 4:
        cases.append(("Staggered", lambda: gb.ImprovedStaggeredFermion(U, U, grid, 0.1), 1, 1146.0))
    for name, make, Ls, fps in cases:
        D = make()
        mk = (lambda: gb.LatticeStaggeredFermion(grid, 1, gb.F32, gb.HALF)) if name == "Staggered" else (lambda: gb.LatticeFermion(grid, Ls, gb.F32, gb.HALF))
        src, out = mk().random(2), mk()
        src.set_checkerboard(gb.Odd)
        ms = timed(lambda: D.DhopEO(src, out, 0), ncall)
        gf = fps * V4 * Ls / 2 / ms / 1e6
        if rank == 0:
            ref = BOOSTER_PER_GPU.get(name, {}).get(L)
            print(json.dumps({"bench": "usqcd", "op": name + " DhopEO fp32", "local_L": L, "Ls": Ls, "n_gpus": world, "ms": ms, "GFlops_per_gpu": gf,
                              "booster_a100_GFlops_per_gpu": ref, "ratio_to_booster_a100": (gf / ref if ref else None)}), flush=True)
        del D, src, out
    # stream triad on a 5D fp32 field (ref :203-262: z = a x - y, bytes = 3 x field)
    x, y, z = (gb.LatticeFermion(grid, 12, gb.F32).random(s) for s in (3, 4, 5))
    ms = timed(lambda: gb.axpy(z, 0.5, x, y), ncall)
    if rank == 0:
        print(json.dumps({"bench": "usqcd", "op": "stream triad fp32", "local_L": L, "Ls": 12, "bytes": 3 * V4 * 12 * 96, "ms": ms, "GBs_per_gpu": 3 * V4 * 12 * 96 / ms / 1e6}), flush=True)
    del x, y, z, U
if dist:
    dist.destroy_process_group()
