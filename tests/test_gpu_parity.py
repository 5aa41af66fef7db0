"""CUDA path (through the C ABI) against the oracle on the same seeded inputs.

f64 instantiation: bit-exact decisions, iteration counts, flags, counters, histograms and
a-posteriori sums for every reference binary / macro set (sum-product's sums within 1e-9 relative:
CUDA's tanh/log are not glibc's).  f32 instantiation: identical decisions on frames that converge
in both, a-posteriori sums within 1e-5 of the frame's largest |LLR| (north-star tolerance)."""
import os

import numpy as np
import pytest

import cases
from ldpcsimulation_b200 import abi, capi
from oracle.oracle_api import Oracle, code_path, load_codewords

pytestmark = pytest.mark.gpu


def _same(a, b, soft="exact"):
    assert np.array_equal(a.bits, b.bits)
    assert np.array_equal(a.iters, b.iters)
    assert np.array_equal(a.errors, b.errors)
    assert np.array_equal(a.flags, b.flags)
    assert a.counters == b.counters
    assert np.array_equal(a.error_weight_hist, b.error_weight_hist)
    assert np.array_equal(a.iter_hist, b.iter_hist)
    assert np.array_equal(a.phase_hist, b.phase_hist)
    if soft == "exact":
        assert np.array_equal(a.soft, b.soft)
    elif soft:
        np.testing.assert_allclose(a.soft, b.soft, rtol=soft, atol=1e-12)


@pytest.mark.parametrize("code", ["PEG", "802_3_H"])
@pytest.mark.parametrize("variant", [v for v in cases.VARIANTS if v != "NGDBFhw"])
def test_f64_bit_exact(variant, code):
    R, snr = cases.operating_point(variant, code)
    cfg = cases.cfg_for(variant, code=code)
    orc = Oracle(code)
    dec = capi.Decoder(capi.Code(code_path(code)), cfg)
    cws = load_codewords(os.path.join(os.path.dirname(code_path(code)), "data.enc"), 7) if code == "PEG" else None
    F = 12 if cfg.kind != abi.KIND_BP or code == "PEG" else 4
    y, noise, rows, cw = cases.make_inputs(orc.N, cfg, snr, R, F, 777 + len(variant), cws)
    a = orc.decode(cfg, snr, R, y, noise, rows, cw)
    b = dec.decode(snr, R, y, noise, rows, cw)
    kind = cfg.kind
    _same(a, b, soft=(1e-9 if kind == abi.KIND_BP else ("exact" if kind in (abi.KIND_MINSUM, abi.KIND_DDBMP) else None)))


@pytest.mark.parametrize("code", ["802_3", "4000", "4376"])
@pytest.mark.parametrize("variant", ["decodeMinSum", "decodeOffsetMinSum", "decodeSMNGDBF", "decodeDDBMP"])
def test_f64_other_codes(variant, code):
    R, snr = cases.operating_point(variant, code)
    cfg = cases.cfg_for(variant, code=code)
    orc = Oracle(code)
    dec = capi.Decoder(capi.Code(code_path(code)), cfg)
    cws = load_codewords(os.path.join(os.path.dirname(code_path(code)), "data.enc"), 3) if code == "4000" else None
    y, noise, rows, cw = cases.make_inputs(orc.N, cfg, snr, R, 5, 31, cws)
    _same(orc.decode(cfg, snr, R, y, noise, rows, cw), dec.decode(snr, R, y, noise, rows, cw),
          soft="exact" if cfg.kind != abi.KIND_GDBF else None)


@pytest.mark.parametrize("maxphase", [1, 3])
def test_ngdbfhw_bit_exact(maxphase):
    cfg = cases.cfg_for("NGDBFhw", maxphase=maxphase, num_iterations=100)
    orc = Oracle("802_3_H")
    dec = capi.Decoder(capi.Code(code_path("802_3_H")), cfg)
    rng = np.random.default_rng(5)
    y, noise, rows, cw = cases.make_inputs(orc.N, cfg, 4.0, 0.8413, 24, 7)
    qp = rng.integers(0, 600, size=24).astype(np.int32)          # any window start the reference could carry in
    _same(orc.decode(cfg, 4.0, 0.8413, y, noise, qpointer0=qp), dec.decode(4.0, 0.8413, y, noise, qpointer0=qp), soft=None)
    # with random (non-)codewords: the uncodedErrors quirk (c in {0,1})
    cws = rng.integers(0, 2, size=(24, orc.N)).astype(np.uint8)
    y2 = (1.0 - 2.0 * cws) * np.abs(y)
    _same(orc.decode(cfg, 4.0, 0.8413, y2, noise, codeword=cws), dec.decode(4.0, 0.8413, y2, noise, codeword=cws), soft=None)


def test_exact_ties_and_zeros():
    rng = np.random.default_rng(8)
    for variant in ("decodeMinSum", "decodeOffsetMinSum", "decodeDDBMP", "decodeGDBF", "decodeSGDBF"):
        cfg = cases.cfg_for(variant, Ymax=1.5, Q=3, delta=0.5)
        orc = Oracle("PEG")
        dec = capi.Decoder(capi.Code(code_path("PEG")), cfg)
        y = rng.integers(-3, 4, size=(8, orc.N)).astype(np.float64) * 0.5
        _same(orc.decode(cfg, 3.0, 0.5, y), dec.decode(3.0, 0.5, y), soft="exact" if cfg.kind != abi.KIND_GDBF else None)


def test_t_extremes_and_ragged_batches():
    orc = Oracle("PEG")
    code = capi.Code(code_path("PEG"))
    for variant, T, F in [("decodeMinSum", 0, 3), ("decodeMinSum", 1, 1), ("decodeMinSum", 50, 5), ("decodeBP", 1, 2),
                          ("decodeDDBMP", 0, 2), ("decodeSMNGDBF", 3, 7), ("decodeSMGDBF", 5, 33)]:
        cfg = cases.cfg_for(variant, num_iterations=T)
        R, snr = cases.operating_point(variant, "PEG")
        y, noise, rows, cw = cases.make_inputs(orc.N, cfg, snr, R, F, 5)
        soft = "exact" if cfg.kind in (abi.KIND_MINSUM, abi.KIND_DDBMP) else (1e-9 if cfg.kind == abi.KIND_BP else None)
        _same(orc.decode(cfg, snr, R, y, noise, rows, cw), capi.Decoder(code, cfg).decode(snr, R, y, noise, rows, cw), soft=soft)
    # empty batch
    cfg = cases.cfg_for("decodeMinSum")
    out = capi.Decoder(code, cfg).decode(2.0, 0.5, np.zeros((0, orc.N)))
    assert out.counters["totalWords"] == 0


@pytest.mark.parametrize("variant,code", [("decodeMinSum", "PEG"), ("decodeNormalizedMinSum", "802_3_H"),
                                          ("decodeOffsetMinSum", "802_3_H"), ("decodeBP", "PEG"), ("decodeBP", "802_3_H")])
def test_f32_within_tolerance(variant, code):
    """fp32 instantiation vs the double oracle.  Tolerance: a-posteriori sums within 1e-5 of the frame's
    largest |LLR| on frames the oracle converges on (non-converging min-sum trajectories are chaotic,
    SURVEY.md 7.2 hard part 2); decisions identical on those frames."""
    R, snr = cases.operating_point(variant, code)
    snr += 0.8                                               # mostly-converging operating point
    cfg64 = cases.cfg_for(variant, code=code)
    cfg32 = cases.cfg_for(variant, code=code, precision=abi.PREC_F32)
    orc = Oracle(code)
    dec = capi.Decoder(capi.Code(code_path(code)), cfg32)
    y, noise, rows, cw = cases.make_inputs(orc.N, cfg64, snr, R, 48, 4242)
    a = orc.decode(cfg64, snr, R, y)
    b = dec.decode(snr, R, y)
    conv = a.errors == 0
    assert conv.sum() >= 24
    assert np.array_equal(a.iters, b.iters)
    assert np.array_equal(a.bits[conv], b.bits[conv])
    scale = np.abs(a.soft).max(axis=1, keepdims=True)
    rel = np.abs(a.soft - b.soft) / scale
    tol = 1e-5 if cfg64.kind == abi.KIND_MINSUM else 2e-5
    assert rel[conv].max() < tol, rel[conv].max()
    # f32 samples in, f32 sums out (the host mains' fast path) decode the same way
    c = dec.decode(snr, R, y.astype(np.float32), y_dtype=abi.DT_F32)
    assert np.array_equal(b.bits[conv], c.bits[conv]) or np.mean(b.bits[conv] != c.bits[conv]) < 1e-3


# ---- config 4: codes whose per-frame state exceeds one SM -> HBM-resident instantiations ----------
@pytest.mark.parametrize("variant,prec", [("decodeMinSum", abi.PREC_F64), ("decodeNormalizedMinSum", abi.PREC_F64),
                                          ("decodeSMNGDBF", abi.PREC_F64), ("decodeMinSum", abi.PREC_F32)])
def test_dvbs2_hbm_resident(variant, prec):
    """DVB-S2 rate-1/2 (N=64800, E=226799): messages live in an HBM workspace; same results as the oracle."""
    cfg = cases.cfg_for(variant, num_iterations=12, precision=prec, alpha=(1.25 if "MinSum" in variant else 0.7))
    orc = Oracle("dvbs2")
    dec = capi.Decoder(capi.Code(code_path("dvbs2")), cfg)
    assert dec.geometry()["smem_bytes"] < 160 * 1024           # messages are not in shared memory (only the tile's packed decisions)
    snr = 1.6 if cfg.kind == abi.KIND_MINSUM else 3.5
    y, noise, rows, cw = cases.make_inputs(orc.N, cfg, snr, 0.5, 3, 17)
    cfg64 = cases.cfg_for(variant, num_iterations=12, alpha=(1.25 if "MinSum" in variant else 0.7))
    a = orc.decode(cfg64, snr, 0.5, y, noise, rows, cw)
    b = dec.decode(snr, 0.5, y, noise, rows, cw)
    if prec == abi.PREC_F64:
        _same(a, b, soft="exact" if cfg.kind == abi.KIND_MINSUM else None)
    else:
        assert np.array_equal(a.iters, b.iters)
        scale = np.abs(a.soft).max(axis=1, keepdims=True)
        assert (np.abs(a.soft - b.soft) / scale).max() < 1e-4
        assert np.mean(a.d != b.d) < 1e-3


@pytest.mark.parametrize("variant", ["decodeMinSum", "decodeBP", "decodeDDBMP", "decodeSMNGDBF", "decodeSGDBF"])
def test_forced_hbm_state_matches_shared_memory_path(variant, monkeypatch):
    """The HBM-resident instantiation is the same code over a different address space."""
    cfg = cases.cfg_for(variant)
    R, snr = cases.operating_point(variant, "PEG")
    orc = Oracle("PEG")
    y, noise, rows, cw = cases.make_inputs(orc.N, cfg, snr, R, 9, 23)
    code = capi.Code(code_path("PEG"))
    a = capi.Decoder(code, cfg).decode(snr, R, y, noise, rows, cw)
    monkeypatch.setenv("LDPC_GPU_FORCE_HBM_STATE", "1")
    dec = capi.Decoder(code, cfg)
    b = dec.decode(snr, R, y, noise, rows, cw)
    _same(a, b, soft="exact" if cfg.kind != abi.KIND_GDBF else None)
    _same(orc.decode(cfg, snr, R, y, noise, rows, cw), b,
          soft=("exact" if cfg.kind in (abi.KIND_MINSUM, abi.KIND_DDBMP) else (1e-9 if cfg.kind == abi.KIND_BP else None)))


@pytest.mark.parametrize("prec", [abi.PREC_F64, abi.PREC_F32])
@pytest.mark.parametrize("variant", ["decodeNormalizedMinSum", "decodeOffsetMinSum", "decodeMinSum", "decodeSaturatedMinSum"])
def test_register_resident_c2v_kernel_equals_scheduled_kernel(variant, prec, monkeypatch):
    """ms_rc_kernel (c2v in the row thread's registers, variable phase publishes only the a-posteriori sum,
    integer-pattern select) computes v2c = sum - c2v with the operands and order of ms_sched_kernel:
    decisions, iteration counts, counters AND a-posteriori sums are bit-identical, in fp32 as in fp64."""
    R, snr = 0.8413, 3.8
    code = capi.Code(code_path("802_3_H"))
    cws = code.random_codewords(5, 4)
    for T in (0, 1, 2, 9):
        cfg = cases.cfg_for(variant, code="802_3_H", num_iterations=T, precision=prec)
        y, noise, rows, cw = cases.make_inputs(2048, cfg, snr, R, 37, 1000 + T, cws if T != 1 else None)
        monkeypatch.delenv("LDPC_GPU_NO_RC", raising=False)
        rc = capi.Decoder(code, cfg)
        assert rc.geometry()["smem_bytes"] == 16 + (4 if prec == abi.PREC_F32 else 8) * 2048 * 9 + 8 * 64 + 16
        a = rc.decode(snr, R, y, codeword=cw)
        monkeypatch.setenv("LDPC_GPU_NO_RC", "1")
        b = capi.Decoder(code, cfg).decode(snr, R, y, codeword=cw)
        assert np.array_equal(a.bits, b.bits) and np.array_equal(a.iters, b.iters) and np.array_equal(a.flags, b.flags)
        assert np.array_equal(a.soft, b.soft) and a.counters == b.counters


@pytest.mark.parametrize("variant", ["decodeNormalizedMinSum", "decodeOffsetMinSum", "decodeMinSum"])
def test_f16x2_register_resident_kernel_equals_f16x2_scheduled_kernel(variant, monkeypatch):
    """ms_h2rc_kernel reorganises ms_h2_kernel (c2v pairs in registers, sums published by the variable phase,
    xorsign minima, integer-pattern select) without changing a single binary16 operation: bit-identical outputs."""
    R, snr = 0.8413, 3.9
    code = capi.Code(code_path("802_3_H"))
    cws = code.random_codewords(9, 4)
    monkeypatch.setenv("LDPC_GPU_NO_X2", "1")             # the labelled kernels, also where the exact-lattice kernel would be selected
    for T, F in ((0, 6), (1, 7), (10, 41)):
        cfg = cases.cfg_for(variant, code="802_3_H", num_iterations=T, precision=abi.PREC_F16X2)
        y, noise, rows, cw = cases.make_inputs(2048, cfg, snr, R, F, 77 + T, cws if T else None)
        monkeypatch.delenv("LDPC_GPU_NO_RC", raising=False)
        a = capi.Decoder(code, cfg).decode(snr, R, y, codeword=cw)
        monkeypatch.setenv("LDPC_GPU_NO_RC", "1")
        b = capi.Decoder(code, cfg).decode(snr, R, y, codeword=cw)
        assert np.array_equal(a.bits, b.bits) and np.array_equal(a.soft, b.soft) and a.counters == b.counters


@pytest.mark.parametrize("prec", [abi.PREC_F64, abi.PREC_F32, abi.PREC_F16X2])
@pytest.mark.parametrize("variant", ["decodeNormalizedMinSum", "decodeOffsetMinSum"])
def test_quantiser_level_input_equals_raw_sample_input(variant, prec):
    """LDPC_GPU_DT_Q8: the samples as a Q-bit converter delivers them (signed quantiser levels, one byte each).  The
    decoder sees exactly what quantize() (src/decodeMinSum.cpp:480-489) makes of the raw double samples: same decisions,
    iteration counts, counters and a-posteriori sums, in every instantiation."""
    R, snr = 0.8413, 3.9
    code = capi.Code(code_path("802_3_H"))
    cfg = cases.cfg_for(variant, code="802_3_H", num_iterations=7, precision=prec)
    dec = capi.Decoder(code, cfg)
    cws = code.random_codewords(3, 4)
    y, noise, rows, cw = cases.make_inputs(2048, cfg, snr, R, 29, 5150, cws)
    y[0, :8] = [0.0, -0.0, 1e-9, -1e-9, cfg.Ymax, -cfg.Ymax, cfg.Ymax * (1 + 1e-12), -3 * cfg.Ymax]      # zero level, saturation edge
    k = abi.quantizer_levels(y, cfg.Ymax, cfg.Q)
    assert k.dtype == np.int8 and np.abs(k).min() == 1 and np.abs(k).max() == 32
    a = dec.decode(snr, R, y, codeword=cw)
    b = dec.decode(snr, R, k, codeword=cw, y_dtype=abi.DT_Q8)
    assert np.array_equal(a.bits, b.bits) and np.array_equal(a.iters, b.iters) and a.counters == b.counters
    assert b.soft.dtype == np.float32 and np.array_equal(a.soft.astype(np.float32), b.soft)


def test_quantiser_level_input_needs_a_quantising_min_sum_decoder():
    code = capi.Code(code_path("802_3_H"))
    dec = capi.Decoder(code, cases.cfg_for("decodeMinSum", code="802_3_H"))
    with pytest.raises(capi.LdpcGpuError):
        dec.decode(4.0, 0.8413, np.ones((2, 2048), np.int8), y_dtype=abi.DT_Q8)


def test_fp16_sample_input_equals_fp32_input_of_the_same_values():
    """LDPC_GPU_DT_F16 only changes how the samples travel: binary16 -> fp32 is exact."""
    cfg = cases.cfg_for("decodeNormalizedMinSum", precision=abi.PREC_F32)
    dec = capi.Decoder(capi.Code(code_path("802_3_H")), cfg)
    y, _, _, _ = cases.make_inputs(2048, cfg, 4.0, 0.8413, 40, 9)
    y16 = y.astype(np.float16)
    a = dec.decode(4.0, 0.8413, y16.astype(np.float32), y_dtype=abi.DT_F32)
    b = dec.decode(4.0, 0.8413, y16, y_dtype=abi.DT_F16)
    assert np.array_equal(a.bits, b.bits) and np.array_equal(a.soft, b.soft) and a.counters == b.counters
    assert b.soft.dtype == np.float32


@pytest.mark.parametrize("variant", ["decodeNormalizedMinSum", "decodeOffsetMinSum", "decodeMinSum"])
def test_f16x2_decisions_on_converged_frames(variant):
    """LDPC_GPU_PREC_F16X2 (two frames per thread, binary16 messages, clamped) is a labelled throughput
    instantiation, not the reference's arithmetic.  What it must keep: the decisions of the double oracle
    on frames that converge, for even and odd frame counts, with and without codewords."""
    R, snr = 0.8413, 4.6
    cfg64 = cases.cfg_for(variant, code="802_3_H", num_iterations=10)
    cfgh = cases.cfg_for(variant, code="802_3_H", num_iterations=10, precision=abi.PREC_F16X2)
    orc = Oracle("802_3_H")
    code = capi.Code(code_path("802_3_H"))
    dec = capi.Decoder(code, cfgh)
    assert dec.geometry()["ctas_per_sm"] >= 2
    cws = code.random_codewords(1, 8)
    for F, use_cw in ((64, False), (33, True)):
        y, noise, rows, cw = cases.make_inputs(orc.N, cfg64, snr, R, F, 99 + F, cws if use_cw else None)
        a = orc.decode(cfg64, snr, R, y, codeword=cw)
        b = dec.decode(snr, R, y, codeword=cw)
        conv = a.errors == 0
        assert conv.sum() >= 0.5 * F
        assert np.array_equal(a.bits[conv], b.bits[conv])
        assert np.array_equal(a.iters, b.iters)
        assert b.counters["totalWords"] == F and b.counters["uncodedErrors"] == a.counters["uncodedErrors"]
        rel = np.abs(np.sign(a.soft[conv]) - np.sign(b.soft[conv])).max()
        assert rel == 0


def test_f16x2_refuses_other_codes_and_decoders():
    with pytest.raises(capi.LdpcGpuError) as e:
        capi.Decoder(capi.Code(code_path("PEG")), cases.cfg_for("decodeMinSum", precision=abi.PREC_F16X2))
    assert e.value.code == abi.ERR_UNSUPPORTED
    with pytest.raises(capi.LdpcGpuError) as e:
        capi.Decoder(capi.Code(code_path("802_3_H")), cases.cfg_for("decodeBP", precision=abi.PREC_F16X2))
    assert e.value.code == abi.ERR_UNSUPPORTED


@pytest.mark.parametrize("variant", ["decodeMinSum", "decodeNormalizedMinSum", "decodeOffsetMinSum", "decodeSaturatedMinSum"])
def test_small_code_multi_frame_kernel_equals_single_frame_kernel(variant, monkeypatch):
    """ms_quad_kernel (small codes: four frames per thread in float4 words, c2v in the row thread's registers, sums
    published by the variable phase) performs the fp32 operations of ms_fast_kernel<float> in the same order: decisions,
    iteration counts, flags, counters and a-posteriori sums bit-identical on every frame -- ragged tiles (F not a multiple
    of the tile), codewords and T = 0 included."""
    R, snr = 0.5, 2.2
    code = capi.Code(code_path("PEG"))
    cws = code.random_codewords(11, 3)
    for T, F in ((0, 5), (1, 9), (7, 8), (50, 43)):
        cfg = cases.cfg_for(variant, code="PEG", num_iterations=T, precision=abi.PREC_F32)
        y, noise, rows, cw = cases.make_inputs(code.N, cfg, snr, R, F, 4400 + T, cws if T != 1 else None)
        monkeypatch.delenv("LDPC_GPU_NO_QUAD", raising=False)
        q = capi.Decoder(code, cfg)
        assert q.geometry()["smem_bytes"] > 64 * 1024                      # the multi-frame kernel was selected
        a = q.decode(snr, R, y, codeword=cw)
        monkeypatch.setenv("LDPC_GPU_NO_QUAD", "1")
        b = capi.Decoder(code, cfg).decode(snr, R, y, codeword=cw)
        assert np.array_equal(a.bits, b.bits) and np.array_equal(a.iters, b.iters) and np.array_equal(a.flags, b.flags)
        assert np.array_equal(a.soft, b.soft) and a.counters == b.counters
        sa = q.simulate(snr, R, 5, 100, 1001).counters                     # Philox entry, ragged last tile
        monkeypatch.setenv("LDPC_GPU_NO_QUAD", "1")
        sb = capi.Decoder(code, cfg).simulate(snr, R, 5, 100, 1001).counters
        assert sa == sb
