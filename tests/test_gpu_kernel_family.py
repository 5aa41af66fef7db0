"""Every kernel that can serve a (decoder, code) pair computes the same thing.

pick_kernel (csrc/ldpc_gpu.cu) chooses the most specialised kernel a code admits; the environment switches below force the
more general kernel of each family.  All of them perform the reference's operations in the reference's order, so their outputs
must be IDENTICAL -- decisions, iteration counts, flags, counters and a-posteriori sums -- in fp64 and in fp32 alike (the fp64
path is the one pinned to the oracle in test_gpu_parity.py; this file pins the other kernels to it and to each other)."""
import numpy as np
import pytest

import cases
from ldpcsimulation_b200 import abi, capi
from oracle.oracle_api import code_path

pytestmark = pytest.mark.gpu

MS_SWITCHES = [(), ("LDPC_GPU_NO_QUAD",), ("LDPC_GPU_NO_QUAD", "LDPC_GPU_NO_SMALL"), ("LDPC_GPU_NO_RC",), ("LDPC_GPU_NO_SCHED",),
               ("LDPC_GPU_GENERIC_MS",), ("LDPC_GPU_FORCE_HBM_STATE",), ("LDPC_GPU_FORCE_HBM_STATE", "LDPC_GPU_NO_TILE")]
ALL = sorted({s for sw in MS_SWITCHES for s in sw} | {"LDPC_GPU_GENERIC_GDBF", "LDPC_GPU_GENERIC_BP"})


def _decode(monkeypatch, switches, code, cfg, snr, R, y, cw, noise=None, rows=0):
    for s in ALL:
        monkeypatch.delenv(s, raising=False)
    for s in switches:
        monkeypatch.setenv(s, "1")
    dec = capi.Decoder(code, cfg)
    geo = dec.geometry()
    out = dec.decode(snr, R, y, codeword=cw, noise=noise, noise_rows=rows) if noise is not None else dec.decode(snr, R, y, codeword=cw)
    return out, (geo["block"], geo["smem_bytes"], geo["ctas_per_sm"])


@pytest.mark.parametrize("prec", [abi.PREC_F64, abi.PREC_F32])
@pytest.mark.parametrize("code_name", ["PEG", "802_3_H", "802_3", "4000", "4376"])
@pytest.mark.parametrize("variant", ["decodeMinSum", "decodeOffsetMinSum"])
def test_min_sum_kernels_agree(variant, code_name, prec, monkeypatch):
    R, snr = cases.operating_point(variant, code_name)
    code = capi.Code(code_path(code_name))
    cws = code.random_codewords(17, 2)
    cfg = cases.cfg_for(variant, code=code_name, num_iterations=5, precision=prec)
    y, _, _, cw = cases.make_inputs(code.N, cfg, snr, R, 21, 900, cws)
    ref, geos = None, set()
    for sw in MS_SWITCHES:
        out, geo = _decode(monkeypatch, sw, code, cfg, snr, R, y, cw)
        geos.add(geo)
        if ref is None:
            ref = out
            continue
        assert np.array_equal(ref.bits, out.bits) and np.array_equal(ref.iters, out.iters) and np.array_equal(ref.flags, out.flags), sw
        assert np.array_equal(ref.soft, out.soft) and ref.counters == out.counters, sw
    assert len(geos) >= 2                                  # the switches really selected different kernels


@pytest.mark.parametrize("code_name", ["PEG", "802_3_H"])
@pytest.mark.parametrize("variant", ["decodeGDBF", "decodeMNGDBF", "decodeSMNGDBF", "decodeRSMNGDBF"])
def test_bit_flipping_kernels_agree(variant, code_name, monkeypatch):
    """gdbf_par_kernel (bit-packed decisions, incremental syndromes) against gdbf_kernel (full recomputation), fp64."""
    R, snr = cases.operating_point(variant, code_name)
    code = capi.Code(code_path(code_name))
    cfg = cases.cfg_for(variant, code=code_name)
    y, noise, rows, cw = cases.make_inputs(code.N, cfg, snr, R, 13, 31)
    a, ga = _decode(monkeypatch, (), code, cfg, snr, R, y, cw, noise, rows)
    b, gb = _decode(monkeypatch, ("LDPC_GPU_GENERIC_GDBF",), code, cfg, snr, R, y, cw, noise, rows)
    assert np.array_equal(a.bits, b.bits) and np.array_equal(a.iters, b.iters) and np.array_equal(a.flags, b.flags)
    assert a.counters == b.counters


def test_sum_product_fp32_kernels_agree_within_tolerance(monkeypatch):
    """fp32 sum-product: the O(dc) phi-domain row update on the register-resident / scheduled structure against the generic
    mp_kernel (same phi, leave-one-out by re-summation).  Not the same operation order, so: identical decisions on frames
    that converge, sums within 2e-5 of the frame's largest |LLR|."""
    R, snr = 0.8413, 4.0
    code = capi.Code(code_path("802_3_H"))
    cfg = cases.cfg_for("decodeBP", code="802_3_H", num_iterations=8, precision=abi.PREC_F32)
    y, _, _, cw = cases.make_inputs(code.N, cfg, snr, R, 33, 5)
    a, _ = _decode(monkeypatch, (), code, cfg, snr, R, y, cw)
    for sw in (("LDPC_GPU_NO_RC",), ("LDPC_GPU_GENERIC_BP",)):
        b, _ = _decode(monkeypatch, sw, code, cfg, snr, R, y, cw)
        ok = a.errors == 0
        assert ok.sum() > 20 and np.array_equal(a.bits[ok], b.bits[ok])
        rel = np.abs(a.soft[ok].astype(np.float64) - b.soft[ok]).max(axis=1) / np.abs(a.soft[ok]).max(axis=1)
        assert rel.max() < 2e-5, sw
