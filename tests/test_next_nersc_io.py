"""SURVEY 8(f) row 4 -- NERSC gauge-configuration I/O (ref: Grid/parallelIO/NerscIO.h:63-290, MetaData.h:143-215; driver
tests/IO/Test_nersc_io.cc), so that real ensembles can enter through LatticeGaugeField import.

The reader / writer is host code of the product library and runs without a GPU, so the interoperability is proven here on
the CPU in BOTH directions against the compiled reference: files written by the reference's NerscIO::writeConfiguration
(3x3 and two-row) are read back bit-exactly with its checksum / plaquette / link trace, and files written by
gb_nersc_write_host pass the reference's own NerscIO::readConfiguration QA.  Committed fixture: a 4^4 two-row file written
by the reference (tests/golden/nersc_4x4x4x4_2row.cfg, 37 kB).  The device wrappers (gb_gauge_read/write_nersc) are the GPU test.
"""
import os

import numpy as np
import pytest

import grid_b200 as gb
from grid_b200 import synthetic as syn
from oracle import pyref as pr

HERE = os.path.dirname(os.path.abspath(__file__))
FIXTURE = os.path.join(HERE, "golden", "nersc_4x4x4x4_2row.cfg")
G = np.load(os.path.join(HERE, "golden", "dirac_golden.npz"))
DIMS = (4, 4, 4, 4)


def plaquette_numpy(U, dims):
    """mean plaquette, the slow and obvious way (ref: WilsonLoops::avgPlaquette, WilsonLoops.h:121-126)"""
    L = dims
    Ug = U.reshape(L[3], L[2], L[1], L[0], 4, 3, 3)
    tot = 0.0
    for mu in range(4):
        for nu in range(mu):
            a = Ug[..., mu, :, :]
            b = np.roll(Ug[..., nu, :, :], -1, axis=3 - mu)
            c = np.roll(Ug[..., mu, :, :], -1, axis=3 - nu)
            d = Ug[..., nu, :, :]
            tot += np.einsum("...ij,...jk,...lk,...il->...", a, b, np.conj(c), np.conj(d)).sum().real
    return tot / U.shape[0] / 6.0 / 3.0


def test_committed_reference_file_reads_back_exactly():
    U, h = gb.NerscIO.read_host(FIXTURE)
    assert tuple(h.dimension) == DIMS and h.data_type == b"4D_SU3_GAUGE" and h.floating_point == b"IEEE64BIG"
    assert h.computed_checksum == h.checksum
    assert abs(h.computed_plaquette - h.plaquette) < 1e-9 and abs(h.computed_link_trace - h.link_trace) < 1e-9
    assert abs(h.plaquette - plaquette_numpy(G["U"], DIMS)) < 1e-9
    # two-row storage: rows 0, 1 are the stored bits, row 2 is reconstructed (unitarity to rounding)
    assert np.array_equal(U[:, :, :2, :], G["U"][:, :, :2, :])
    assert np.max(np.abs(U - G["U"])) < 1e-14


def test_write_then_read_roundtrip_and_header_fields(tmp_path):
    U = syn.hot_gauge((4, 6, 4, 8), seed=9)
    for two_row in (0, 1):
        path = tmp_path / f"cfg{two_row}"
        gb.NerscIO.write_host(path, U, (4, 6, 4, 8), two_row, ens_label="lbl", ens_id="idx", sequence_number=42)
        V, h = gb.NerscIO.read_host(path)
        assert tuple(h.dimension) == (4, 6, 4, 8) and h.sequence_number == 42 and h.ensemble_label == b"lbl" and h.ensemble_id == b"idx"
        assert h.data_type == (b"4D_SU3_GAUGE" if two_row else b"4D_SU3_GAUGE_3x3")
        assert abs(h.plaquette - plaquette_numpy(U, (4, 6, 4, 8))) < 1e-9
        assert abs(h.link_trace - np.einsum("smii->", U).real / U.shape[0] / 12.0) < 1e-9
        assert np.array_equal(V, U) if not two_row else np.max(np.abs(V - U)) < 1e-14
        assert os.path.getsize(path) == h.data_start + U.shape[0] * 4 * (2 if two_row else 3) * 3 * 16


def test_corrupt_files_are_rejected(tmp_path):
    """checksum exact, plaquette 1e-5, link trace 1e-6 (ref: NerscIO.h:196-210); a flipped payload bit or an edited header must fail"""
    U = syn.hot_gauge(DIMS, seed=10)
    path = tmp_path / "cfg"
    gb.NerscIO.write_host(path, U, DIMS)
    raw = bytearray(open(path, "rb").read())
    start = gb.NerscIO.readHeader(path).data_start
    bad = bytearray(raw); bad[start + 1000] ^= 0x10
    open(tmp_path / "flip", "wb").write(bad)
    with pytest.raises(gb.GridB200Error, match="checksum"):
        gb.NerscIO.read_host(tmp_path / "flip")
    txt = bytes(raw[:start]).decode()
    plaq = [l for l in txt.splitlines() if l.startswith("PLAQUETTE")][0]
    edited = txt.replace(plaq, "PLAQUETTE  = 0.5").encode()
    open(tmp_path / "plaq", "wb").write(edited + bytes(raw[start:]))
    with pytest.raises(gb.GridB200Error, match="plaquette"):
        gb.NerscIO.read_host(tmp_path / "plaq")
    open(tmp_path / "short", "wb").write(bytes(raw[:-64]))
    with pytest.raises(gb.GridB200Error, match="shorter"):
        gb.NerscIO.read_host(tmp_path / "short")
    open(tmp_path / "nothdr", "wb").write(b"hello\n")
    with pytest.raises(gb.GridB200Error, match="BEGIN_HEADER"):
        gb.NerscIO.read_host(tmp_path / "nothdr")


def test_ieee32big_payload(tmp_path):
    """single-precision archives (FLOATING_POINT = IEEE32BIG) as the reference reads them (NerscIO.h:163-167)"""
    U = syn.hot_gauge(DIMS, seed=11)
    path = tmp_path / "cfg64"
    gb.NerscIO.write_host(path, U, DIMS)
    h = gb.NerscIO.readHeader(path)
    U32 = U.astype(np.complex64)
    payload = U32.view(np.float32).astype(">f4").tobytes()
    words = np.frombuffer(payload, dtype=">u4").astype(np.uint64)
    csum = int(words.sum() % (1 << 32))
    hdr = open(path, "rb").read()[: h.data_start].decode()
    hdr = hdr.replace("IEEE64BIG", "IEEE32BIG")
    cs_line = [l for l in hdr.splitlines() if l.startswith("CHECKSUM")][0]
    hdr = hdr.replace(cs_line, f"CHECKSUM = {csum:x}")
    open(tmp_path / "cfg32", "wb").write(hdr.encode() + payload)
    V, h32 = gb.NerscIO.read_host(tmp_path / "cfg32")
    assert h32.floating_point == b"IEEE32BIG" and h32.computed_checksum == csum
    assert np.array_equal(V, U32.astype(np.complex128))


@pytest.mark.skipif(not pr.available(), reason="oracle/_ref/libgridref.so not built (needs /root/reference)")
@pytest.mark.parametrize("two_row", [0, 1])
def test_interoperates_with_the_reference_both_ways(tmp_path, two_row):
    dims = (4, 4, 8, 4)
    U = syn.hot_gauge(dims, seed=12)
    # reference writes -> we read
    p1 = tmp_path / "by_reference"
    pr.nersc_write(dims, U, p1, two_row)
    V, h = gb.NerscIO.read_host(p1)
    assert h.computed_checksum == h.checksum and h.sequence_number == 7
    assert np.array_equal(V[:, :, :2, :], U[:, :, :2, :]) and np.max(np.abs(V - U)) < 1e-14
    # we write -> the reference reads and runs its own QA (exits on checksum mismatch, asserts on plaquette / link trace)
    p2 = tmp_path / "by_gridb200"
    gb.NerscIO.write_host(p2, U, dims, two_row)
    W, plaq, link = pr.nersc_read(dims, p2)
    assert np.max(np.abs(W - U)) < 1e-14
    assert abs(plaq - h.plaquette) < 1e-9 and abs(link - h.link_trace) < 1e-9
    # same payload bytes as the reference's file
    a, b = open(p1, "rb").read(), open(p2, "rb").read()
    assert a[h.data_start:] == b[gb.NerscIO.readHeader(p2).data_start:]


@pytest.mark.gpu
def test_device_gauge_field_read_write(tmp_path):
    ctx = gb.Context(0)
    grid = gb.GridCartesian(ctx, DIMS)
    Umu = gb.LatticeGaugeField(grid, gb.F64)
    h = gb.NerscIO.readConfiguration(Umu, FIXTURE)
    assert h.computed_checksum == h.checksum
    assert np.max(np.abs(Umu.export_lex() - G["U"])) < 1e-14
    gb.NerscIO.writeConfiguration(Umu, tmp_path / "out", two_row=0)
    V, h2 = gb.NerscIO.read_host(tmp_path / "out")
    assert np.max(np.abs(V - G["U"])) < 1e-14 and abs(h2.plaquette - h.plaquette) < 1e-9
    # a single-precision operator takes the same file
    Uf = gb.LatticeGaugeField(grid, gb.F32)
    gb.NerscIO.readConfiguration(Uf, FIXTURE)
    assert np.max(np.abs(Uf.export_lex() - G["U"])) < 1e-6
