"""LDPC_GPU_PREC_F16X2 on an exact lattice (csrc/ldpc_ms_x2.cuh): two frames per lane in binary16, decisions CERTIFIED
identical to the reference's unbounded doubles, fp64 redo of the frames the kernel cannot certify.

Criterion: bit-exact decisions, iteration counts, error counts and counters against the double oracle on EVERY frame,
converged or not; a-posteriori sums equal wherever the c2v cap has not engaged (|sum| below the cap)."""
import os

import numpy as np
import pytest

import cases
from ldpcsimulation_b200 import abi, capi
from oracle.oracle_api import Oracle, code_path

pytestmark = pytest.mark.gpu

R = 0.8413
OMS = dict(flags=["quantizeSamples", "offsetMS"], num_iterations=10, Ymax=1.9375, Q=5, delta=0.125)
MS = dict(flags=["quantizeSamples"], num_iterations=10, Ymax=1.9375, Q=5)
CAP = (2047 - 31) // 6 / 16.0                       # c2v cap in LLR units for Q = 5, dv = 6


def _dec(kw, **over):
    code = capi.Code(code_path("802_3_H"))
    d = capi.Decoder(code, abi.default_cfg(abi.KIND_MINSUM, precision=abi.PREC_F16X2, **dict(kw, **over)))
    return d


def _check(a, b, soft=True):
    assert np.array_equal(a.bits, b.bits)
    assert np.array_equal(a.iters, b.iters) and np.array_equal(a.errors, b.errors) and np.array_equal(a.flags, b.flags)
    assert a.counters == b.counters and np.array_equal(a.error_weight_hist, b.error_weight_hist)
    if soft:                                          # sums are exact while no c2v of the variable has been capped
        small = np.abs(a.soft) < CAP
        assert np.array_equal(a.soft[small], b.soft[small].astype(np.float64))


@pytest.mark.parametrize("snr,kw", [(3.6, OMS), (4.0, OMS), (4.0, MS), (5.0, OMS)])
def test_x2_decisions_equal_oracle_on_all_frames(snr, kw):
    F = 2049                                          # odd: the last pair has a dead lane
    orc = Oracle("802_3_H")
    cfg64 = abi.default_cfg(abi.KIND_MINSUM, **kw)
    y, _, _, _ = cases.make_inputs(orc.N, cfg64, snr, R, F, 4000 + int(snr * 10))
    a = orc.decode(cfg64, snr, R, y)
    dec = _dec(kw)
    assert dec.stats()[0], "the exact-lattice kernel was not selected"
    b = dec.decode(snr, R, y)
    _check(a, b)
    if snr < 4.5:
        assert (a.errors > 0).sum() >= 10             # non-converged frames are part of the comparison
    # quantiser levels in (the e2e format) and fp32 samples in
    c = dec.decode(snr, R, abi.quantizer_levels(y, 1.9375, 5), y_dtype=abi.DT_Q8)
    assert np.array_equal(a.bits, c.bits) and a.counters == c.counters
    print("redo frames:", dec.stats()[1], "of", 2 * F)


def test_x2_with_codewords_and_T_extremes():
    orc = Oracle("802_3_H")
    code = capi.Code(code_path("802_3_H"))
    cws = code.random_codewords(5, 16)
    for T in (0, 1, 2, 25):
        kw = dict(OMS, num_iterations=T)
        cfg64 = abi.default_cfg(abi.KIND_MINSUM, **kw)
        y, _, _, cw = cases.make_inputs(orc.N, cfg64, 3.9, R, 300, 17 + T, cws)
        a = orc.decode(cfg64, 3.9, R, y, codeword=cw)
        b = capi.Decoder(code, abi.default_cfg(abi.KIND_MINSUM, precision=abi.PREC_F16X2, **kw)).decode(3.9, R, y, codeword=cw)
        _check(a, b, soft=(T <= 2))
    out = capi.Decoder(code, abi.default_cfg(abi.KIND_MINSUM, precision=abi.PREC_F16X2, **OMS)).decode(4.0, R, np.zeros((0, orc.N)))
    assert out.counters["totalWords"] == 0


def test_x2_redo_path_is_exact(monkeypatch):
    """A cap just above the stable-state bound engages long before most frames are stable: many frames go through the fp64
    redo launch, and the results are still the oracle's."""
    monkeypatch.setenv("LDPC_GPU_X2_CAP_UNITS", "12")
    orc = Oracle("802_3_H")
    cfg64 = abi.default_cfg(abi.KIND_MINSUM, **OMS)
    y, _, _, _ = cases.make_inputs(orc.N, cfg64, 3.8, R, 1500, 99)
    a = orc.decode(cfg64, 3.8, R, y)
    dec = _dec(OMS)
    b = dec.decode(3.8, R, y)
    redo = dec.stats()[1]
    assert redo > 20, redo
    assert np.array_equal(a.bits, b.bits) and np.array_equal(a.errors, b.errors) and a.counters == b.counters
    assert np.array_equal(a.iters, b.iters) and np.array_equal(a.flags, b.flags)
    # with codewords (the host pipeline collects the uncertified frames of all chunks and re-decodes them, compacted, at the end)
    cws = capi.Code(code_path("802_3_H")).random_codewords(3, 8)
    yc, _, _, cw = cases.make_inputs(orc.N, cfg64, 3.8, R, 1201, 100, cws)
    ac, bc = orc.decode(cfg64, 3.8, R, yc, codeword=cw), dec.decode(3.8, R, yc, codeword=cw)
    assert dec.stats()[1] > redo
    assert np.array_equal(ac.bits, bc.bits) and np.array_equal(ac.errors, bc.errors) and np.array_equal(ac.iters, bc.iters) and ac.counters == bc.counters
    # the throughput entry goes through the same redo launch
    s16 = dec.simulate(3.8, R, 77, 1000, 6000).counters
    monkeypatch.delenv("LDPC_GPU_X2_CAP_UNITS")
    s64 = capi.Decoder(capi.Code(code_path("802_3_H")), abi.default_cfg(abi.KIND_MINSUM, **OMS)).simulate(3.8, R, 77, 1000, 6000).counters
    assert s16 == s64


@pytest.mark.parametrize("snr", [3.4, 4.0, 4.6])
def test_x2_simulate_counters_equal_fp64_instantiation(snr):
    """Same Philox frames through the packed kernel and through the fp64 parity instantiation: every counter and histogram
    equal (the fp64 instantiation is pinned to the oracle by test_gpu_parity / test_gpu_simulate)."""
    code = capi.Code(code_path("802_3_H"))
    F = 60001
    d16 = capi.Decoder(code, abi.default_cfg(abi.KIND_MINSUM, precision=abi.PREC_F16X2, **OMS))
    d64 = capi.Decoder(code, abi.default_cfg(abi.KIND_MINSUM, **OMS))
    a, b = d16.simulate(snr, R, 2026, 12345, F), d64.simulate(snr, R, 2026, 12345, F)
    assert a.counters == b.counters and np.array_equal(a.error_weight_hist, b.error_weight_hist)
    assert a.counters["totalWords"] == F
    print("snr", snr, "redo", d16.stats()[1], "word errors", a.counters["wordErrors"])
    o = Oracle("802_3_H").simulate(abi.default_cfg(abi.KIND_MINSUM, **OMS), snr, R, 2026, 12345, 64).counters
    assert d16.simulate(snr, R, 2026, 12345, 64).counters == o


def test_x2_certified_stop_reports_the_same_results():
    """LDPC_GPU_F_CERT_STOP: iterations after a frame's decisions are certified final are skipped; decisions, iteration counts
    (T, as the reference accounts them) and counters are still those of T full iterations."""
    orc = Oracle("802_3_H")
    cfg64 = abi.default_cfg(abi.KIND_MINSUM, **OMS)
    stop = dict(OMS, flags=OMS["flags"] + ["certStop"])
    for snr in (3.7, 4.4):
        y, _, _, _ = cases.make_inputs(orc.N, cfg64, snr, R, 777, 31)
        a = orc.decode(cfg64, snr, R, y)
        b = _dec(stop).decode(snr, R, y, want_soft=False)
        assert np.array_equal(a.bits, b.bits) and np.array_equal(a.iters, b.iters) and np.array_equal(a.errors, b.errors)
        assert a.counters == b.counters
    d = _dec(stop)
    s1 = d.simulate(4.0, R, 9, 0, 50000).counters
    s2 = capi.Decoder(capi.Code(code_path("802_3_H")), abi.default_cfg(abi.KIND_MINSUM, **OMS)).simulate(4.0, R, 9, 0, 50000).counters
    assert s1 == s2
    with pytest.raises(capi.LdpcGpuError):                                   # sums of iteration T are not available
        d.decode(4.0, R, np.ones((2, orc.N)))
    with pytest.raises(capi.LdpcGpuError):                                   # and the flag exists for the exact-lattice kernel only
        capi.Decoder(capi.Code(code_path("802_3_H")), abi.default_cfg(abi.KIND_MINSUM, precision=abi.PREC_F32, **stop))


@pytest.mark.parametrize("prec", [abi.PREC_F16X2, abi.PREC_F32, abi.PREC_F64])
def test_bit_packed_levels_equal_byte_levels(prec):
    """LDPC_GPU_DT_QP: the quantiser levels packed Q bits per sample decode exactly like the one-byte levels and the raw samples."""
    orc = Oracle("802_3_H")
    code = capi.Code(code_path("802_3_H"))
    for kw, Ymax, Q in ((OMS, 1.9375, 5), (dict(flags=["quantizeSamples", "normalizedMS"], num_iterations=7, Ymax=2.0, Q=6, alpha=1.25), 2.0, 6)):
        if prec == abi.PREC_F16X2 and Q == 6:
            continue
        cfg = abi.default_cfg(abi.KIND_MINSUM, precision=prec, **kw)
        y, _, _, _ = cases.make_inputs(orc.N, cfg, 3.9, R, 333, 8)
        y[0, :6] = [0.0, -0.0, 1e-9, Ymax, -Ymax * (1 + 1e-12), 5.0]
        dec = capi.Decoder(code, cfg)
        a = dec.decode(3.9, R, abi.quantizer_levels(y, Ymax, Q), y_dtype=abi.DT_Q8)
        b = dec.decode(3.9, R, abi.quantizer_levels_packed(y, Ymax, Q), y_dtype=abi.DT_QP)
        c = dec.decode(3.9, R, y)
        assert np.array_equal(a.bits, b.bits) and np.array_equal(a.soft, b.soft) and a.counters == b.counters
        assert np.array_equal(a.bits, c.bits) and a.counters == c.counters


def test_non_lattice_configs_keep_the_labelled_kernel():
    d = _dec(dict(flags=["quantizeSamples", "normalizedMS"], num_iterations=10, Ymax=2.0, Q=6, alpha=1.25))
    assert not d.stats()[0]
    d = _dec(dict(flags=["quantizeSamples", "offsetMS"], num_iterations=10, Ymax=2.0, Q=5, delta=0.125))   # step 4/31: not dyadic
    assert not d.stats()[0]


def test_fast_channel_mode():
    """LDPC_GPU_CHANNEL_FAST: SFU Box-Muller.  The dumped samples are what the decoder sees (dump -> oracle decode == simulate
    counters), and they are standard normal to Monte-Carlo accuracy."""
    code = capi.Code(code_path("802_3_H"))
    orc = Oracle("802_3_H")
    cfg64 = abi.default_cfg(abi.KIND_MINSUM, channel_mode=abi.CHANNEL_FAST, **OMS)
    for prec in (abi.PREC_F16X2, abi.PREC_F32, abi.PREC_F64):
        dec = capi.Decoder(code, abi.default_cfg(abi.KIND_MINSUM, precision=prec, channel_mode=abi.CHANNEL_FAST, **OMS))
        y, _ = dec.channel_dump(4.0, R, 5, 100, 512)
        ref = orc.decode(cfg64, 4.0, R, y).counters
        assert dec.simulate(4.0, R, 5, 100, 512).counters == ref
    sigma = cases.sigma_of(4.0, R)
    n = (y - 1.0) / sigma
    assert abs(n.mean()) < 5e-3 and abs(n.var() - 1.0) < 5e-3
    assert abs((n ** 4).mean() - 3.0) < 5e-2 and np.abs(n).max() < 6.7
    # and the two generators are the same Philox stream through two Box-Muller evaluations: samples agree to ~1e-5
    y0, _ = capi.Decoder(code, abi.default_cfg(abi.KIND_MINSUM, **OMS)).channel_dump(4.0, R, 5, 100, 512)
    assert np.abs(y - y0).max() < 1e-4
