"""LDPC_GPU_PREC_F16X2 on codes whose messages live in HBM (csrc/ldpc_ms_tileh.cuh): binary16 message tiles of 64 frames on the
exact lattice, fp64 redo of the frames in which a check-to-variable magnitude left the range binary16 holds exactly.

Criterion: bit-exact decisions, iteration counts, error counts, flags, counters AND a-posteriori sums against the double oracle on
every frame (no cap is applied to the messages: a frame either stays exact or is re-decoded in fp64)."""
import numpy as np
import pytest

import cases
from ldpcsimulation_b200 import abi, capi
from oracle.oracle_api import Oracle, code_path

pytestmark = pytest.mark.gpu

OMS = dict(flags=["quantizeSamples", "offsetMS"], num_iterations=10, Ymax=1.9375, Q=5, delta=0.125)
MS = dict(flags=["quantizeSamples"], num_iterations=10, Ymax=1.9375, Q=5)


def _same(a, b, soft=True):
    assert np.array_equal(a.bits, b.bits)
    assert np.array_equal(a.iters, b.iters) and np.array_equal(a.errors, b.errors) and np.array_equal(a.flags, b.flags)
    assert a.counters == b.counters and np.array_equal(a.error_weight_hist, b.error_weight_hist)
    if soft:
        assert np.array_equal(a.soft, b.soft.astype(np.float64))


def _tile_decoder(name, kw):
    d = capi.Decoder(capi.Code(code_path(name)), abi.default_cfg(abi.KIND_MINSUM, precision=abi.PREC_F16X2, **kw))
    assert d.stats()[0], "the exact-lattice kernel was not selected"
    assert d.geometry()["smem_bytes"] < 227 * 1024         # only the cp.async staging slots are in shared memory, not the messages (DVB-S2: 37.8 MB per tile)
    return d


@pytest.mark.parametrize("kw,snr", [(OMS, 1.4), (OMS, 2.2), (MS, 1.8)])
def test_dvbs2_binary16_tiles_equal_oracle(kw, snr):
    """DVB-S2 rate 1/2 (N = 64800, E = 226799): 67 frames = one full tile of 64 and a ragged one."""
    orc = Oracle("dvbs2")
    cfg64 = abi.default_cfg(abi.KIND_MINSUM, **kw)
    y, _, _, _ = cases.make_inputs(orc.N, cfg64, snr, 0.5, 67, 300 + int(10 * snr))
    a = orc.decode(cfg64, snr, 0.5, y)
    dec = _tile_decoder("dvbs2", kw)
    b = dec.decode(snr, 0.5, y)
    _same(a, b)
    assert (a.errors > 0).sum() >= 3                       # frames that have not converged are part of the comparison
    assert dec.stats()[1] == 0                             # no frame needed the fp64 decoder at this operating point
    c = dec.decode(snr, 0.5, abi.quantizer_levels_packed(y, 1.9375, 5), y_dtype=abi.DT_QP)     # the e2e sample format
    assert np.array_equal(a.bits, c.bits) and a.counters == c.counters


def test_dvbs2_binary16_tiles_with_codewords_and_T_extremes():
    orc = Oracle("dvbs2")
    code = capi.Code(code_path("dvbs2"))
    # (the dense GF(2) encoder is limited to M*N <= 2^28 bits; the decoder and the error counts take any transmitted bit pattern)
    cws = np.random.default_rng(16).integers(0, 2, (3, orc.N), dtype=np.uint8)
    for T in (0, 1, 3):
        kw = dict(OMS, num_iterations=T)
        cfg64 = abi.default_cfg(abi.KIND_MINSUM, **kw)
        y, _, _, cw = cases.make_inputs(orc.N, cfg64, 1.8, 0.5, 9, 40 + T, cws)
        a = orc.decode(cfg64, 1.8, 0.5, y, codeword=cw)
        b = capi.Decoder(code, abi.default_cfg(abi.KIND_MINSUM, precision=abi.PREC_F16X2, **kw)).decode(1.8, 0.5, y, codeword=cw)
        _same(a, b)
    out = _tile_decoder("dvbs2", OMS).decode(2.0, 0.5, np.zeros((0, orc.N)))
    assert out.counters["totalWords"] == 0


def test_dvbs2_redo_path_is_exact(monkeypatch):
    """A low cap sends many frames through the fp64 redo launch (frame-list indirection of ms_tile_kernel<double>): results are
    still the oracle's, through the parity entry and through the throughput entry."""
    monkeypatch.setenv("LDPC_GPU_X2_CAP_UNITS", "40")      # |c2v| > 2.5
    orc = Oracle("dvbs2")
    cfg64 = abi.default_cfg(abi.KIND_MINSUM, **OMS)
    y, _, _, _ = cases.make_inputs(orc.N, cfg64, 2.6, 0.5, 70, 7)
    a = orc.decode(cfg64, 2.6, 0.5, y)
    dec = _tile_decoder("dvbs2", OMS)
    b = dec.decode(2.6, 0.5, y)
    redo = dec.stats()[1]
    assert 0 < redo, redo
    _same(a, b)
    s16 = dec.simulate(2.6, 0.5, 77, 1000, 130).counters
    assert dec.stats()[1] > redo
    monkeypatch.delenv("LDPC_GPU_X2_CAP_UNITS")
    s64 = capi.Decoder(capi.Code(code_path("dvbs2")), cfg64).simulate(2.6, 0.5, 77, 1000, 130).counters
    assert s16 == s64
    print("redo frames:", redo, "of 70")


def test_dvbs2_simulate_counters_equal_fp64_instantiation():
    code = capi.Code(code_path("dvbs2"))
    d16 = capi.Decoder(code, abi.default_cfg(abi.KIND_MINSUM, precision=abi.PREC_F16X2, **OMS))
    d64 = capi.Decoder(code, abi.default_cfg(abi.KIND_MINSUM, **OMS))
    F = 64 * 148 + 33                                      # more tiles than CTAs, ragged
    a, b = d16.simulate(2.0, 0.5, 2026, 555, F), d64.simulate(2.0, 0.5, 2026, 555, F)
    assert a.counters == b.counters and np.array_equal(a.error_weight_hist, b.error_weight_hist)
    assert a.counters["totalWords"] == F
    print("redo", d16.stats()[1], "word errors", a.counters["wordErrors"])


def test_small_code_forced_into_hbm_tiles(monkeypatch):
    """16-bit index instantiation: the (3,6) PEG code forced onto the HBM path, irregular-free but ragged tiles."""
    monkeypatch.setenv("LDPC_GPU_FORCE_HBM_STATE", "1")
    orc = Oracle("PEG")
    cfg64 = abi.default_cfg(abi.KIND_MINSUM, **OMS)
    y, _, _, _ = cases.make_inputs(orc.N, cfg64, 2.0, 0.5, 333, 5)
    a = orc.decode(cfg64, 2.0, 0.5, y)
    dec = _tile_decoder("PEG", OMS)
    b = dec.decode(2.0, 0.5, y)
    _same(a, b)
    print("redo frames:", dec.stats()[1], "of 333")
