"""gb_op_dhop_host on N GPUs (torchrun, one rank per GPU): per-site parity of the pipelined decomposed form against the CPU oracle on
a small global lattice, then its time at local 32^4 x Ls16 with pinned host buffers, pipelined and as import + hop + export.
   python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port 29544 scripts/mgpu_host_dhop.py"""
import os, sys, json, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
import torch.distributed as dist
import grid_b200 as gb
from grid_b200 import synthetic as syn, decomp
from oracle import pyoracle as po

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


def site_err(a, b):
    a = a.reshape(a.shape[0], -1).astype(np.complex128); b = b.reshape(b.shape[0], -1).astype(np.complex128)
    return float(np.max(np.linalg.norm(a - b, axis=1) / np.maximum(np.linalg.norm(b, axis=1), 1e-300)))


fails = 0
for local, Ls, kind in (((8, 8, 8, 8), 16, "dwf"), ((16, 4, 6, 4), 8, "mobius")):
    gd = tuple(l * m for l, m in zip(local, mpi))
    U = syn.hot_gauge(gd, seed=3); src = syn.random_fermion(gd, Ls, seed=4)
    orc = po.OracleOp(1, gd, Ls, mass=0.1, M5=1.8, b=1.5 if kind == "mobius" else 1.0, c=0.5 if kind == "mobius" else 0.0, prec=1)
    orc.import_gauge(U)
    grid = gb.GridCartesian(ctx, gd, mpi)
    for prec, tol in ((gb.F32, 1e-6), (gb.F64, 1e-13)):
        Umu = gb.LatticeGaugeField(grid, prec).import_lex(decomp.scatter(U, gd, mpi, rank))
        D = gb.DomainWallFermion(Umu, grid, Ls, 0.1, 1.8) if kind == "dwf" else gb.MobiusFermion(Umu, grid, Ls, 0.1, 1.8, 1.5, 0.5)
        h = decomp.scatter(src, gd, mpi, rank, inner=Ls).astype(gb._cdtype(prec))
        for dag in (0, 1):
            ref = decomp.scatter(orc.apply(po.OP_DHOP, src, dag=dag), gd, mpi, rank, inner=Ls)
            for call in range(3):        # both epoch parities of the receive buffers
                e = mx(site_err(D.Dhop_host(h, np.empty_like(h), dag), ref))
                ok = e < tol; fails += not ok
                if rank == 0:
                    print(f"{'ok  ' if ok else 'FAIL'} mpi {mpi} {kind} Ls{Ls} prec{prec} Dhop_host dag{dag} call{call}: {e:.3e}", flush=True)
        # compressed halos (fp32 operator: bf16 on the wire, fp64 operator: fp32) on real peers: within the comms precision, every form
        fin = gb.LatticeFermion(grid, Ls, prec).import_lex(h); out = gb.LatticeFermion(grid, Ls, prec)
        D.set_halo_compression(True)
        ref = decomp.scatter(orc.apply(po.OP_DHOP, src, dag=0), gd, mpi, rank, inner=Ls)
        for overlap in (1, 2, 0):
            D.set_overlap(overlap)
            D.Dhop(fin, out, 0)
            e = mx(site_err(out.export_lex(), ref)); lo, hi = (20 * tol, 8e-3 if prec == gb.F32 else 2e-6)
            ok = lo < e < hi; fails += not ok
            if rank == 0:
                print(f"{'ok  ' if ok else 'FAIL'} mpi {mpi} {kind} Ls{Ls} prec{prec} overlap{overlap} Dhop compressed halos: {e:.3e} (expected in ({lo:.0e}, {hi:.0e}))", flush=True)
        e = mx(site_err(D.Dhop_host(h, np.empty_like(h), 0), ref)); ok = e < hi; fails += not ok
        if rank == 0:
            print(f"{'ok  ' if ok else 'FAIL'} mpi {mpi} {kind} Ls{Ls} prec{prec} Dhop_host compressed halos: {e:.3e}", flush=True)
        D.set_halo_compression(False); D.set_overlap(1)
        D.Dhop(fin, out, 0)
        e = mx(site_err(out.export_lex(), ref)); ok = e < tol; fails += not ok
        if rank == 0:
            print(f"{'ok  ' if ok else 'FAIL'} mpi {mpi} {kind} Ls{Ls} prec{prec} Dhop after compression switched off: {e:.3e}", flush=True)
local, Ls = (32, 32, 32, 32), 16
gd = [l * m for l, m in zip(local, mpi)]
grid = gb.GridCartesian(ctx, gd, mpi)
D = gb.DomainWallFermion(gb.LatticeGaugeField(grid, gb.F32).random(1), grid, Ls, 0.1, 1.8)
srcf = gb.LatticeFermion(grid, Ls, gb.F32).random(2); out = gb.LatticeFermion(grid, Ls, gb.F32)
D.Dhop(srcf, out, 0)
ref = out.export_lex()
n = srcf.local_sites
hin = torch.empty((n, 4, 3), dtype=torch.complex64).pin_memory().numpy(); hout = torch.empty((n, 4, 3), dtype=torch.complex64).pin_memory().numpy()
hin[...] = srcf.export_lex()
for env in ("1", "0"):
    os.environ["GB_HOST_PIPE_DECOMP"] = env
    D.Dhop_host(hin, hout, 0); D.Dhop_host(hin, hout, 0)
    err = mx(float(np.max(np.abs(hout - ref)) / np.max(np.abs(ref))))
    ctx.synchronize(); dist.barrier(); ctx.synchronize()
    t0 = time.perf_counter()
    for _ in range(6):
        D.Dhop_host(hin, hout, 0)
    ctx.synchronize()
    s = mx((time.perf_counter() - t0) / 6)
    if rank == 0:
        print(json.dumps({"n_gpus": world, "mpi": mpi, "local": local, "Ls": Ls, "pipelined_decomposed": env == "1", "ms_per_call": s * 1e3,
                          "GBs_per_direction_per_gpu": hin.nbytes / s / 1e9, "max_err_vs_device_hop": err}), flush=True)
if rank == 0:
    print("MGPU_HOST_DHOP " + ("PASS" if fails == 0 else f"FAIL ({fails})"), flush=True)
dist.barrier()
dist.destroy_process_group()
sys.exit(1 if fails else 0)
