"""Host-side logic of the N>1 path, on CPU with the gloo backend (world size 2): rank <-> processor-coordinate map,
neighbour tables (library vs python), scatter/gather of lexicographic fields, and the face exchange pattern the halo
code uses (my x_mu=0 slice feeds the backward neighbour's forward leg; my x_mu=L-1 slice feeds the forward neighbour's
backward leg) checked against global periodic neighbours on a coordinate-encoded field (cf. tests/Test_stencil.cc:70-131)."""
import os

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

import grid_b200 as gb
from grid_b200 import decomp


def _free_port():
    """a port nobody listens on right now (the rendezvous of each test gets its own)"""
    import socket
    with socket.socket(socket.AF_INET, socket.SOCK_STREAM) as so:
        so.bind(("127.0.0.1", 0))
        return so.getsockname()[1]


def encode(gdims):
    v = int(np.prod(gdims))
    i = np.arange(v)
    x = i % gdims[0]; y = (i // gdims[0]) % gdims[1]; z = (i // (gdims[0] * gdims[1])) % gdims[2]; t = i // (gdims[0] * gdims[1] * gdims[2])
    return np.stack([x, y, z, t], axis=1).astype(np.int64)


def test_geometry_matches_python_rule():
    for gdims, mpi in (((8, 8, 8, 16), (1, 1, 2, 4)), ((8, 4, 4, 4), (2, 1, 1, 1)), ((8, 8, 8, 8), (2, 2, 2, 1))):
        n = int(np.prod(mpi))
        seen = set()
        for rank in range(n):
            ld, origin, nbr = gb.geometry_query(gdims, mpi, rank)
            assert ld == decomp.local_dims(gdims, mpi)
            pc = decomp.rank_to_pcoor(rank, mpi)
            assert origin == tuple(p * l for p, l in zip(pc, ld))
            assert nbr == decomp.neighbour_ranks(rank, mpi)
            assert decomp.pcoor_to_rank(pc, mpi) == rank
            seen.add(origin)
        assert len(seen) == n
    with pytest.raises(gb.GridB200Error):
        gb.geometry_query((8, 8, 8, 9), (1, 1, 1, 2), 0)


def test_scatter_gather_roundtrip_single_process():
    gdims, mpi, Ls = (4, 4, 8, 8), (1, 1, 2, 2), 3
    rng = np.random.default_rng(0)
    f = rng.standard_normal((int(np.prod(gdims)) * Ls, 4, 3)) + 0j
    parts = [decomp.scatter(f, gdims, mpi, r, inner=Ls) for r in range(4)]
    assert np.array_equal(decomp.gather(parts, gdims, mpi, inner=Ls), f)
    # every global site appears exactly once
    enc = encode(gdims)
    allc = np.concatenate([decomp.scatter(enc, gdims, mpi, r) for r in range(4)])
    assert len({tuple(c) for c in allc}) == len(enc)


def _worker(rank, world, port, mpi, gdims, q):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        ld, origin, nbr = gb.geometry_query(gdims, mpi, rank)
        enc = encode(gdims)
        loc = decomp.scatter(enc, gdims, mpi, rank).reshape(ld[3], ld[2], ld[1], ld[0], 4)
        ok = True
        for mu in range(4):
            if mpi[mu] == 1:
                continue
            ax = 3 - mu  # array axis of dimension mu
            lo = np.ascontiguousarray(np.take(loc, 0, axis=ax))            # x_mu = 0 slice  -> backward neighbour
            hi = np.ascontiguousarray(np.take(loc, ld[mu] - 1, axis=ax))   # x_mu = L-1 slice -> forward neighbour
            fwd, bwd = nbr[mu]
            recv_f, recv_b = torch.empty_like(torch.from_numpy(lo)), torch.empty_like(torch.from_numpy(hi))
            ops = [dist.P2POp(dist.isend, torch.from_numpy(lo), bwd), dist.P2POp(dist.irecv, recv_f, fwd),
                   dist.P2POp(dist.isend, torch.from_numpy(hi), fwd), dist.P2POp(dist.irecv, recv_b, bwd)]
            for r in dist.batch_isend_irecv(ops):
                r.wait()
            # what the forward leg of my x_mu = L-1 sites must see: global neighbour x+mu (periodic)
            want_f = hi.copy(); want_f[..., mu] = (want_f[..., mu] + 1) % gdims[mu]
            want_b = lo.copy(); want_b[..., mu] = (want_b[..., mu] - 1) % gdims[mu]
            ok = ok and np.array_equal(recv_f.numpy(), want_f) and np.array_equal(recv_b.numpy(), want_b)
        # global sum over ranks (GlobalSum analogue)
        t = torch.tensor([float(loc.sum())], dtype=torch.float64)
        dist.all_reduce(t)
        ok = ok and t.item() == float(enc.sum())
        gathered = [None] * world
        dist.all_gather_object(gathered, decomp.scatter(enc, gdims, mpi, rank))
        if rank == 0:
            ok = ok and np.array_equal(decomp.gather(gathered, gdims, mpi), enc)
        q.put((rank, bool(ok)))
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("mpi,gdims", [((1, 1, 1, 2), (4, 4, 4, 8)), ((2, 1, 1, 1), (8, 4, 4, 4)), ((1, 1, 2, 1), (4, 4, 8, 4))])
def test_face_exchange_pattern_world2_gloo(mpi, gdims):
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, mpi, gdims, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = [q.get(timeout=120) for _ in procs]
    for p in procs:
        p.join(timeout=60)
    assert sorted(res) == [(0, True), (1, True)]


def _naik_worker(rank, world, port, mpi, gdims, q):
    """Three-deep halos of the improved staggered operator (grid_b200/csrc/stag.cu: stag_exchange): my slices 0..2 fill the
    BACKWARD neighbour's forward halo (its x_mu = L..L+2), my slices L-3..L-1 the FORWARD neighbour's backward halo
    (its x_mu = -3..-1).  With the slabs attached, every +-1 and +-3 neighbour of every local site is the global periodic one."""
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        ld, origin, nbr = gb.geometry_query(gdims, mpi, rank)
        enc = encode(gdims)
        loc = decomp.scatter(enc, gdims, mpi, rank).reshape(ld[3], ld[2], ld[1], ld[0], 4)
        ok = True
        for mu in range(4):
            ax = 3 - mu
            if mpi[mu] == 1:      # undecomposed: periodic wrap inside the local volume
                ext = np.concatenate([np.take(loc, range(ld[mu] - 3, ld[mu]), axis=ax), loc, np.take(loc, range(0, 3), axis=ax)], axis=ax)
            else:
                first = np.ascontiguousarray(np.take(loc, range(0, 3), axis=ax))               # halo dir 0 of my backward neighbour
                last = np.ascontiguousarray(np.take(loc, range(ld[mu] - 3, ld[mu]), axis=ax))   # halo dir 1 of my forward neighbour
                fwd, bwd = nbr[mu]
                recv_f, recv_b = torch.empty_like(torch.from_numpy(first)), torch.empty_like(torch.from_numpy(last))
                ops = [dist.P2POp(dist.isend, torch.from_numpy(first), bwd), dist.P2POp(dist.irecv, recv_f, fwd),
                       dist.P2POp(dist.isend, torch.from_numpy(last), fwd), dist.P2POp(dist.irecv, recv_b, bwd)]
                for r in dist.batch_isend_irecv(ops):
                    r.wait()
                ext = np.concatenate([recv_b.numpy(), loc, recv_f.numpy()], axis=ax)
            for disp in (1, -1, 3, -3):
                got = np.take(ext, range(3 + disp, 3 + disp + ld[mu]), axis=ax)
                want = loc.copy(); want[..., mu] = (want[..., mu] + disp) % gdims[mu]
                ok = ok and np.array_equal(got, want)
        q.put((rank, bool(ok)))
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("mpi,gdims", [((1, 1, 1, 2), (4, 4, 4, 8)), ((2, 1, 1, 1), (8, 4, 4, 4)), ((1, 2, 1, 1), (4, 12, 4, 4))])
def test_naik_three_deep_halo_pattern_world2_gloo(mpi, gdims):
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_naik_worker, args=(r, 2, port, mpi, gdims, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = [q.get(timeout=120) for _ in procs]
    for p in procs:
        p.join(timeout=60)
    assert sorted(res) == [(0, True), (1, True)]
