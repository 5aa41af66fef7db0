"""Parity of the THROUGHPUT instantiations on all frames, converged or not.

* exact lattice: offset min-sum with Ymax = 1.9375, Q = 5, delta = 0.125 puts every channel value, message and sum
  on the grid of multiples of 1/16 (quantize(), src/decodeMinSum.cpp:480-489: step 2*Ymax/31 = 1/8, saturation value
  15.5/8).  As long as |value| * 16 < 2^24 every fp32 operation of the decoder is exact, so the fp32 instantiation must
  equal the double oracle bit for bit: decisions, iteration counts AND a-posteriori sums, on every frame.
* the bench's normalised workload (alpha = 1.25, Ymax = 2, Q = 6: not a dyadic lattice) cannot be bit-exact in fp32;
  the test REPORTS the fraction of frames whose decisions differ and asserts that each of them is a frame the double
  oracle does not converge on (gpurun_out/parity_report_*.json, copied to profiles/ by hand).
* fp32 GDBF family: criterion stated in test_f32_gdbf_family_criterion.
"""
import json
import os

import numpy as np
import pytest

import cases
from ldpcsimulation_b200 import abi, capi
from oracle.oracle_api import Oracle, code_path

pytestmark = pytest.mark.gpu

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LATTICE = dict(flags=["quantizeSamples", "offsetMS"], num_iterations=10, Ymax=1.9375, Q=5, delta=0.125)
R = 0.8413


def _report(name, obj):
    out = os.path.join(ROOT, "gpurun_out")
    os.makedirs(out, exist_ok=True)
    with open(os.path.join(out, "parity_report_%s.json" % name), "w") as f:
        json.dump(obj, f, indent=1, sort_keys=True)


@pytest.mark.parametrize("snr", [3.6, 4.0])
def test_f32_exact_lattice_all_frames(snr):
    F = 2048
    cfg64 = abi.default_cfg(abi.KIND_MINSUM, **LATTICE)
    cfg32 = abi.default_cfg(abi.KIND_MINSUM, precision=abi.PREC_F32, **LATTICE)
    orc = Oracle("802_3_H")
    code = capi.Code(code_path("802_3_H"))
    y, _, _, _ = cases.make_inputs(orc.N, cfg64, snr, R, F, 1000 + int(snr * 10))
    a = orc.decode(cfg64, snr, R, y)
    assert np.abs(a.soft).max() * 16 < 2 ** 24                       # the premise: everything fp32 touches is exact
    assert (a.errors > 0).sum() >= 10                               # non-converged frames are part of the comparison
    b = capi.Decoder(code, cfg32).decode(snr, R, y)
    assert np.array_equal(a.bits, b.bits)
    assert np.array_equal(a.iters, b.iters) and np.array_equal(a.errors, b.errors)
    assert np.array_equal(a.soft, b.soft)
    assert a.counters == b.counters and np.array_equal(a.error_weight_hist, b.error_weight_hist)
    # fp32 samples in / fp32 sums out, and one-byte quantiser levels in: the same frames, the same bits
    c = capi.Decoder(code, cfg32).decode(snr, R, abi.quantizer_levels(y, 1.9375, 5), y_dtype=abi.DT_Q8)
    assert np.array_equal(a.bits, c.bits) and np.array_equal(a.soft.astype(np.float32), c.soft)


@pytest.mark.parametrize("code_name", ["802_3", "4376"])
def test_f32_exact_lattice_other_codes(code_name):
    """The same exactness argument on the full-rank 802.3an H and the (4376,4094) code (whatever kernel serves them)."""
    Rc, snr, _ = cases.CODES[code_name]
    cfg64 = abi.default_cfg(abi.KIND_MINSUM, **LATTICE)
    cfg32 = abi.default_cfg(abi.KIND_MINSUM, precision=abi.PREC_F32, **LATTICE)
    orc = Oracle(code_name)
    y, _, _, _ = cases.make_inputs(orc.N, cfg64, snr + 0.3, Rc, 256, 77)
    a = orc.decode(cfg64, snr + 0.3, Rc, y)
    assert np.abs(a.soft).max() * 16 < 2 ** 24
    b = capi.Decoder(capi.Code(code_path(code_name)), cfg32).decode(snr + 0.3, Rc, y)
    assert np.array_equal(a.bits, b.bits) and np.array_equal(a.soft, b.soft) and a.counters == b.counters


def test_f32_headline_config_mismatch_report():
    """bench.py's default workload (normalised min-sum, Ymax 2, Q 6, alpha 1.25, T = 10, 4.0 dB), fp32 against the double
    oracle on 4096 frames: a number, not a filter."""
    F, snr = 4096, 4.0
    kw = dict(flags=["quantizeSamples", "normalizedMS"], num_iterations=10, Ymax=2.0, Q=6, alpha=1.25)
    cfg64 = abi.default_cfg(abi.KIND_MINSUM, **kw)
    cfg32 = abi.default_cfg(abi.KIND_MINSUM, precision=abi.PREC_F32, **kw)
    orc = Oracle("802_3_H")
    y, _, _, _ = cases.make_inputs(orc.N, cfg64, snr, R, F, 20262)
    a = orc.decode(cfg64, snr, R, y)
    b = capi.Decoder(capi.Code(code_path("802_3_H")), cfg32).decode(snr, R, y)
    differ = np.any(a.bits != b.bits, axis=1)
    conv = a.errors == 0
    scale = np.abs(a.soft).max(axis=1, keepdims=True)
    rel_frame = np.abs(a.soft - b.soft) / scale
    with np.errstate(divide="ignore", invalid="ignore"):
        rel_elem = np.where(a.soft != 0, np.abs(a.soft - b.soft) / np.abs(a.soft), 0.0)
    rep = {
        "frames": F, "snr_db": snr, "oracle_word_errors": int((~conv).sum()),
        "frames_with_any_decision_difference": int(differ.sum()),
        "of_which_converged_in_oracle": int((differ & conv).sum()),
        "bit_decisions_differing": int((np.unpackbits(a.bits ^ b.bits, axis=1)).sum()),
        "fp32_word_errors": int((b.errors > 0).sum()),
        "max_rel_to_frame_max_converged": float(rel_frame[conv].max()),
        "max_rel_per_element_converged": float(rel_elem[conv].max()),
        "p999_rel_per_element_converged": float(np.quantile(rel_elem[conv], 0.999)),
        "max_rel_to_frame_max_nonconverged": float(rel_frame[~conv].max()) if (~conv).any() else 0.0,
    }
    _report("nms_f32", rep)
    assert (differ & conv).sum() == 0, rep                         # every differing frame is a non-converged frame
    assert differ.sum() <= (~conv).sum()
    assert rep["max_rel_to_frame_max_converged"] < 1e-5, rep
    assert np.array_equal(a.iters, b.iters)
    # the two decoders' FER on these frames differ by at most the differing frames
    assert abs(rep["fp32_word_errors"] - rep["oracle_word_errors"]) <= differ.sum()


@pytest.mark.parametrize("variant", ["decodeGDBF", "decodeSMNGDBF", "decodeRSMNGDBF"])
def test_f32_gdbf_family_criterion(variant):
    """fp32 bit-flipping against the double oracle.  A flip is the comparison E_i < theta_i and flips feed back, so fp32
    cannot be decision-exact; the criterion is stated here: on the same samples and noise,
    (1) of the frames the oracle decodes to the codeword, at most 1 % fail in fp32, (2) the fraction of frames whose
    decisions differ is below 5 %, and (3) the word-error counts differ by no more than the differing frames.  fp32 throughput figures are quoted with this criterion, never as parity-exact."""
    code_name = "802_3_H"
    Rc, snr = cases.operating_point(variant, code_name)
    cfg64 = cases.cfg_for(variant, code=code_name)
    cfg32 = cases.cfg_for(variant, code=code_name, precision=abi.PREC_F32)
    orc = Oracle(code_name)
    F = 384
    y, noise, rows, cw = cases.make_inputs(orc.N, cfg64, snr, Rc, F, 99)
    a = orc.decode(cfg64, snr, Rc, y, noise, rows, cw)
    b = capi.Decoder(capi.Code(code_path(code_name)), cfg32).decode(snr, Rc, y, noise, rows, cw)
    differ = np.any(a.bits != b.bits, axis=1)
    easy = a.errors == 0
    rep = {"variant": variant, "frames": F, "differ": int(differ.sum()), "easy": int(easy.sum()),
           "easy_and_fp32_in_error": int((easy & (b.errors > 0)).sum()),
           "oracle_word_errors": int((a.errors > 0).sum()), "fp32_word_errors": int((b.errors > 0).sum()),
           "iters_equal_frames": int((a.iters == b.iters).sum())}
    _report("gdbf_f32_%s" % variant, rep)
    assert rep["easy_and_fp32_in_error"] <= 0.01 * easy.sum(), rep
    assert differ.mean() < 0.05, rep
    assert abs(rep["oracle_word_errors"] - rep["fp32_word_errors"]) <= differ.sum(), rep


@pytest.mark.parametrize("T,snr", [(10, 4.0), (4, 3.4)])
def test_f32_sum_product_error_report(T, snr):
    """fp32 sum-product (phi domain, MUFU ex2 / lg2 / rcp; ms_rc_kernel<.., ALGO_BP>) against the double oracle on the bench's
    code at its operating point (T = 10, 4.0 dB: every LLR ends near the +-MAXLLR clip) and at T = 4, 3.4 dB (unsaturated
    messages, many frames not yet converged).  north_star: identical hard decisions on all but tie frames, LLRs within 1e-5
    relative.  What holds, and is asserted: decisions identical on every frame the oracle converges on; a-posteriori LLRs within
    2e-5 of the frame's largest |LLR| on those frames.  What is REPORTED (gpurun_out/parity_report_bp_f32_*.json): the per-element
    relative error, which is larger where an LLR is the small difference of large messages."""
    F = 512
    cfg64 = abi.default_cfg(abi.KIND_BP, num_iterations=T)
    cfg32 = abi.default_cfg(abi.KIND_BP, num_iterations=T, precision=abi.PREC_F32)
    orc = Oracle("802_3_H")
    y, _, _, _ = cases.make_inputs(orc.N, cfg64, snr, R, F, 4242)
    a = orc.decode(cfg64, snr, R, y)
    b = capi.Decoder(capi.Code(code_path("802_3_H")), cfg32).decode(snr, R, y)
    conv = a.errors == 0
    diff_frames = np.any(a.bits != b.bits, axis=1)
    assert not np.any(diff_frames & conv)
    sa, sb = a.soft[conv], b.soft[conv].astype(np.float64)
    frame_rel = np.abs(sa - sb).max(axis=1) / np.abs(sa).max(axis=1)
    elem_rel = np.abs(sa - sb) / np.maximum(np.abs(sa), 1e-300)
    _report("bp_f32_T%d_%.1fdB" % (T, snr), dict(frames=F, converged=int(conv.sum()), frames_with_different_decisions=int(diff_frames.sum()),
                           different_and_converged=int((diff_frames & conv).sum()),
                           per_frame_rel_err_max=float(frame_rel.max()), per_frame_rel_err_median=float(np.median(frame_rel)),
                           per_element_rel_err_max=float(elem_rel.max()), per_element_rel_err_p999=float(np.quantile(elem_rel, 0.999)),
                           per_element_rel_err_median=float(np.median(elem_rel)),
                           elements_above_1e5=float((elem_rel > 1e-5).mean())))
    assert frame_rel.max() < 2e-5
