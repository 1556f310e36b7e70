"""Host-side domain-decomposition helpers (numpy): which sites a rank owns and how global lexicographic arrays map to
local ones.  Rule (ref: documentation/interfacing.rst:130-146): global site r lives on processor coordinate
r_mu // ldim_mu, local coordinate r_mu % ldim_mu; ranks are lexicographic in the processor grid, dimension 0 fastest.
"""
import numpy as np


def rank_to_pcoor(rank, mpi):
    pc = []
    for m in mpi:
        pc.append(rank % m)
        rank //= m
    return tuple(pc)


def pcoor_to_rank(pc, mpi):
    r, mul = 0, 1
    for c, m in zip(pc, mpi):
        r += mul * (c % m)
        mul *= m
    return r


def local_dims(gdims, mpi):
    assert all(g % m == 0 for g, m in zip(gdims, mpi))
    return tuple(g // m for g, m in zip(gdims, mpi))


def neighbour_ranks(rank, mpi):
    """[(forward, backward)] per dimension, as gb_grid stores them"""
    pc = rank_to_pcoor(rank, mpi)
    out = []
    for d in range(4):
        f = list(pc); f[d] += 1
        b = list(pc); b[d] -= 1
        out.append((pcoor_to_rank(f, mpi), pcoor_to_rank(b, mpi)))
    return out


def _as_grid(a, dims, inner):
    """[V*inner, ...] lexicographic (x fastest, inner fastest of all) -> [t,z,y,x,inner,...]"""
    return a.reshape((dims[3], dims[2], dims[1], dims[0], inner) + a.shape[1:])


def scatter(global_arr, gdims, mpi, rank, inner=1):
    """Local block of a global lexicographic array of site objects ([V*inner, ...])."""
    ld = local_dims(gdims, mpi)
    pc = rank_to_pcoor(rank, mpi)
    g = _as_grid(global_arr, gdims, inner)
    sl = tuple(slice(pc[d] * ld[d], (pc[d] + 1) * ld[d]) for d in (3, 2, 1, 0))
    loc = g[sl]
    return np.ascontiguousarray(loc).reshape((-1,) + global_arr.shape[1:])


def gather(local_arrs, gdims, mpi, inner=1):
    """Inverse of scatter: list of local arrays (indexed by rank) -> global array."""
    ld = local_dims(gdims, mpi)
    tail = local_arrs[0].shape[1:]
    g = np.empty((gdims[3], gdims[2], gdims[1], gdims[0], inner) + tail, dtype=local_arrs[0].dtype)
    for rank, loc in enumerate(local_arrs):
        pc = rank_to_pcoor(rank, mpi)
        sl = tuple(slice(pc[d] * ld[d], (pc[d] + 1) * ld[d]) for d in (3, 2, 1, 0))
        g[sl] = loc.reshape((ld[3], ld[2], ld[1], ld[0], inner) + tail)
    return g.reshape((-1,) + tail)
