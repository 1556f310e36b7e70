"""Host-side synthetic lattice fields in the C-ABI import layout (numpy only).

Restates what the reference benchmarks draw (ref: benchmarks/Benchmark_dwf_fp32.cc:163-190):
 * gauge  : SU<Nc>::HotConfiguration = gaussian -> Ta() -> exp -> SU(3)   (ref: Grid/qcd/utils/GaugeGroup.h:332-349)
 * fermion: random(): real and imaginary parts uniform, then normalised to unit norm for Dhop timing
Grid's per-site sitmo RNG streams are not reproduced; only the distributions are.
Layouts: gauge [V4,4,3,3] complex, fermion [V4*Ls,4,3] complex, lexicographic with x fastest, s fastest of all.
"""
import numpy as np


def random_su3(n, rng, scale=1.0):
    """n random SU(3) matrices exp(Ta(G)), G iid complex gaussian (complex128)."""
    g = scale * (rng.standard_normal((n, 3, 3)) + 1j * rng.standard_normal((n, 3, 3)))
    a = 0.5 * (g - np.conj(np.swapaxes(g, 1, 2)))              # anti-hermitian part
    tr = np.trace(a, axis1=1, axis2=2) / 3.0
    a = a - tr[:, None, None] * np.eye(3)[None]                # traceless
    h = -1j * a                                                # hermitian
    w, v = np.linalg.eigh(h)
    u = (v * np.exp(1j * w)[:, None, :]) @ np.conj(np.swapaxes(v, 1, 2))
    return u


def hot_gauge(dims, seed=1234, dtype=np.complex128, chunk=1 << 18):
    """LatticeGaugeField: [V4,4,3,3]."""
    v4 = int(np.prod(dims))
    rng = np.random.default_rng(seed)
    out = np.empty((v4 * 4, 3, 3), dtype=dtype)
    for i in range(0, v4 * 4, chunk):
        j = min(v4 * 4, i + chunk)
        out[i:j] = random_su3(j - i, rng).astype(dtype)
    return out.reshape(v4, 4, 3, 3)


def unit_gauge(dims, dtype=np.complex128):
    v4 = int(np.prod(dims))
    out = np.zeros((v4, 4, 3, 3), dtype=dtype)
    out[:, :, range(3), range(3)] = 1.0
    return out


def random_fermion(dims, Ls=1, seed=5678, dtype=np.complex128, normalise=False, gaussian=False):
    """LatticeFermion: [V4*Ls,4,3]; uniform [0,1) components like Grid random(), or gaussian."""
    n = int(np.prod(dims)) * Ls
    rng = np.random.default_rng(seed)
    rdt = np.float32 if dtype == np.complex64 else np.float64
    out = np.empty((n, 4, 3), dtype=dtype)
    draw = (lambda: rng.standard_normal((n, 4, 3), dtype=rdt)) if gaussian else (lambda: rng.random((n, 4, 3), dtype=rdt))
    out.real = draw()
    out.imag = draw()
    if normalise:
        out *= 1.0 / np.sqrt(np.vdot(out.astype(np.complex128), out.astype(np.complex128)).real)
    return out
