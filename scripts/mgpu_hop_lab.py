"""Timings of the decomposed fp32 Dhop on N GPUs (torchrun), default form and with the t faces left to the pack kernel.
usage: torchrun ... scripts/mgpu_hop_lab.py  (local volumes 32^4 and 64.64.32.16, Ls 16)"""
import os, sys, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch, torch.distributed as dist
import grid_b200 as gb
rank, world, lrank = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(lrank)
dist.init_process_group("nccl", device_id=torch.device("cuda", lrank))
ctx = gb.Context(lrank)
uid = [gb.Context.unique_id() if rank == 0 else None]
dist.broadcast_object_list(uid, src=0)
ctx.comm_init(rank, world, uid[0])
mpi = {2: (1, 1, 1, 2), 4: (1, 1, 2, 2), 8: (1, 1, 2, 4)}[world]
def mx(x):
    t = torch.tensor([x], dtype=torch.float64, device=f"cuda:{lrank}"); dist.all_reduce(t, op=dist.ReduceOp.MAX); return float(t.item())
for local in ((32, 32, 32, 32), (64, 64, 32, 16)):
    g = [l * m for l, m in zip(local, mpi)]
    grid = gb.GridCartesian(ctx, g, mpi)
    D = gb.DomainWallFermion(gb.LatticeGaugeField(grid, gb.F32).random(1), grid, 16, 0.1, 1.8)
    src = gb.LatticeFermion(grid, 16, gb.F32).random(2); out = gb.LatticeFermion(grid, 16, gb.F32)
    so = gb.LatticeFermion(grid, 16, gb.F32, gb.HALF); ro = gb.LatticeFermion(grid, 16, gb.F32, gb.HALF)
    gb.pickCheckerboard(gb.Odd, so, src)
    for env in ({}, {"GB_HOP_SENDS_T": "0"}):
        os.environ.update(env)
        res = {}
        for name, f in (("Dhop", lambda: D.Dhop(src, out, 0)), ("DhopEO", lambda: D.DhopEO(so, ro, 0))):
            for _ in range(10): f()
            ctx.synchronize(); dist.barrier(); ctx.synchronize()
            l0 = ctx.launch_count(); ctx.timer_start()
            for _ in range(100): f()
            ms = mx(ctx.timer_stop()) / 100
            res[name] = {"ms": round(ms, 4), "launches": (ctx.launch_count() - l0) / 100}
            dist.barrier()
        for k in env: os.environ.pop(k)
        if rank == 0:
            print(json.dumps({"world": world, "mpi": mpi, "local": local, "env": env, **res}), flush=True)
    del D, src, out, so, ro, grid
dist.barrier()
