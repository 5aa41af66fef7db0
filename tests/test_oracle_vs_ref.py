"""The C restatement (oracle/ldpc_oracle.c) against the reference's own object code (oracle/_ref),
bit for bit, on seeded inputs: every reference binary / macro set, several codes, all-zero and
data.enc codewords.  Skipped where oracle/_ref is absent (it is built wherever /root/reference exists
and travels to the GPU box)."""
import os

import numpy as np
import pytest

import cases
from ldpcsimulation_b200 import abi
from oracle.oracle_api import Oracle, Reference, code_path, load_codewords

pytestmark = [pytest.mark.ref, pytest.mark.skipif(not Reference.available(), reason="oracle/_ref not built")]


def _compare(a, b, soft=True):
    assert np.array_equal(a.bits, b.bits)
    assert np.array_equal(a.iters, b.iters)
    assert np.array_equal(a.errors, b.errors)
    assert np.array_equal(a.flags, b.flags)
    assert a.counters == b.counters
    assert np.array_equal(a.error_weight_hist, b.error_weight_hist)
    assert np.array_equal(a.iter_hist, b.iter_hist)
    assert np.array_equal(a.phase_hist, b.phase_hist)
    if soft:
        assert np.array_equal(a.soft, b.soft)          # bit-exact doubles, including BP's tanh/log


@pytest.mark.parametrize("code", ["PEG", "802_3_H"])
@pytest.mark.parametrize("variant", [v for v in cases.VARIANTS if v != "NGDBFhw"])
def test_variant_bit_exact(variant, code):
    R, snr = cases.operating_point(variant, code)
    cfg = cases.cfg_for(variant, code=code)
    orc, ref = Oracle(code), Reference(variant, code)
    assert ref.flags == cfg.flags and ref.kind == cfg.kind
    cws = load_codewords(os.path.join(os.path.dirname(code_path(code)), "data.enc"), 7) if code == "PEG" else None
    F = 10 if cfg.kind != abi.KIND_BP or code == "PEG" else 4
    y, noise, rows, cw = cases.make_inputs(orc.N, cfg, snr, R, F, 20260 + len(variant), cws)
    _compare(orc.decode(cfg, snr, R, y, noise, rows, cw), ref.decode(cfg, snr, R, y, noise, rows, cw))


@pytest.mark.parametrize("code", ["802_3", "4000", "4376"])
@pytest.mark.parametrize("variant", ["decodeMinSum", "decodeNormalizedMinSum", "decodeSMNGDBF", "decodeDDBMP"])
def test_other_codes(variant, code):
    R, snr = cases.operating_point(variant, code)
    cfg = cases.cfg_for(variant, code=code)
    orc, ref = Oracle(code), Reference(variant, code)
    cws = load_codewords(os.path.join(os.path.dirname(code_path(code)), "data.enc"), 3) if code == "4000" else None
    y, noise, rows, cw = cases.make_inputs(orc.N, cfg, snr, R, 4, 99, cws)
    _compare(orc.decode(cfg, snr, R, y, noise, rows, cw), ref.decode(cfg, snr, R, y, noise, rows, cw))


@pytest.mark.parametrize("maxphase", [1, 3])
def test_ngdbfhw_with_carried_qpointer(maxphase):
    """NGDBFhw's noise window position carries from frame to frame (src/NGDBFhw.cpp:356-358); the
    reference's trace of it is replayed into the oracle."""
    cfg = cases.cfg_for("NGDBFhw", maxphase=maxphase, num_iterations=100)
    orc, ref = Oracle("802_3_H"), Reference("NGDBFhw", "802_3_H")
    y, noise, rows, cw = cases.make_inputs(orc.N, cfg, 4.0, 0.8413, 16, 7)
    b = ref.decode(cfg, 4.0, 0.8413, y, noise)
    a = orc.decode(cfg, 4.0, 0.8413, y, noise, qpointer0=b.qpointer_trace[:-1])
    _compare(a, b, soft=False)
    assert b.qpointer_trace.max() > 0


def test_ngdbfhw_codeword_uncoded_quirk():
    """uncodedErrors counts r*c<0 with c in {0,1} (src/NGDBFhw.cpp:141,230): zero for the all-zero word."""
    cfg = cases.cfg_for("NGDBFhw", num_iterations=50)
    orc, ref = Oracle("802_3_H"), Reference("NGDBFhw", "802_3_H")
    rng = np.random.default_rng(3)
    # not codewords of H: the decoder does not care, the accounting quirk is what is pinned here
    cws = rng.integers(0, 2, size=(3, orc.N)).astype(np.uint8)
    y, noise, rows, cw = cases.make_inputs(orc.N, cfg, 4.5, 0.8413, 3, 11, cws)
    qp = np.zeros(3, np.int32)
    a = orc.decode(cfg, 4.5, 0.8413, y, noise, codeword=cw, qpointer0=qp)
    b = ref.decode(cfg, 4.5, 0.8413, y, noise, codeword=cw, qpointer0=qp)
    _compare(a, b, soft=False)
    assert a.counters["uncodedErrors"] > 0


def test_t_extremes():
    """T=1 and a long run; windowsize larger than T (smoothing counts from the first iteration)."""
    for variant, T in [("decodeMinSum", 1), ("decodeMinSum", 50), ("decodeSMNGDBF", 3), ("decodeSMGDBF", 5)]:
        cfg = cases.cfg_for(variant, num_iterations=T, windowsize=64 if "SM" in variant else 0)
        orc, ref = Oracle("PEG"), Reference(variant, "PEG")
        R, snr = cases.operating_point(variant, "PEG")
        y, noise, rows, cw = cases.make_inputs(orc.N, cfg, snr, R, 6, 5)
        _compare(orc.decode(cfg, snr, R, y, noise, rows, cw), ref.decode(cfg, snr, R, y, noise, rows, cw))


def test_exact_ties_and_zeros():
    """Samples on a coarse lattice force exact |v2c| ties and exact zero sums (sgn(0)=+1, d=-1 on sum==0)."""
    rng = np.random.default_rng(8)
    for variant in ("decodeMinSum", "decodeOffsetMinSum", "decodeDDBMP", "decodeGDBF"):
        cfg = cases.cfg_for(variant, Ymax=1.5, Q=3, delta=0.5)
        orc, ref = Oracle("PEG"), Reference(variant, "PEG")
        y = rng.integers(-3, 4, size=(8, orc.N)).astype(np.float64) * 0.5
        _compare(orc.decode(cfg, 3.0, 0.5, y), ref.decode(cfg, 3.0, 0.5, y))
