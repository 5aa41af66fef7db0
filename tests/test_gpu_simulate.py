"""Throughput entry (on-device Philox channel) against the oracle's independent restatement of the
same counter-based channel: raw Philox blocks, dumped samples, and whole-run counters, bit for bit."""
import os

import numpy as np
import pytest

import cases
from ldpcsimulation_b200 import abi, capi
from oracle.oracle_api import Oracle, code_path, load_codewords

pytestmark = pytest.mark.gpu


def test_philox_known_answers_on_device():
    # Random123 kat_vectors, philox4x32-10
    kats = [((0, 0, 0, 0), (0, 0), (0x6627e8d5, 0xe169c58d, 0xbc57ac4c, 0x9b00dbd8)),
            ((0xffffffff,) * 4, (0xffffffff,) * 2, (0x408f276d, 0x41c83b0e, 0xa20bc7c6, 0x6d5451fd)),
            ((0x243f6a88, 0x85a308d3, 0x13198a2e, 0x03707344), (0xa4093822, 0x299f31d0), (0xd16cfe09, 0x94fdcceb, 0x5001e420, 0x24126ea1))]
    for ctr, key, want in kats:
        assert tuple(int(x) for x in capi.philox4x32(ctr, key)) == want


@pytest.mark.parametrize("variant,code", [("decodeMinSum", "PEG"), ("decodeSMNGDBF", "PEG"), ("decodeStochasticNGDBF", "PEG"),
                                          ("decodeUniformMNGDBF", "PEG"), ("NGDBFhw", "802_3_H"), ("decodeRSMNGDBF", "802_3_H")])
def test_channel_dump_matches_oracle(variant, code):
    R, snr = cases.operating_point(variant, code)
    cfg = cases.cfg_for(variant, code=code, num_iterations=6)
    orc = Oracle(code)
    dec = capi.Decoder(capi.Code(code_path(code)), cfg)
    cws = load_codewords(os.path.join(os.path.dirname(code_path(code)), "data.enc"), 5) if code == "PEG" else None
    dec.set_codewords(cws)
    y0, n0 = orc.channel_dump(cfg, snr, R, 1234, 1000, 9, cws)
    y1, n1 = dec.channel_dump(snr, R, 1234, 1000, 9)
    assert np.array_equal(y0, y1)                           # explicitly rounded fp32 Box-Muller: bit-exact CPU vs GPU
    if n0 is not None:
        assert np.array_equal(n0, n1)
    # sanity of the channel itself
    x = 1.0 if cws is None else (1.0 - 2.0 * cws[(1000 + np.arange(9)) % len(cws)])
    n = (y1 / x - 1.0) / cases.sigma_of(snr, R)
    assert abs(n.mean()) < 0.05 and abs(n.std() - 1.0) < 0.05


@pytest.mark.parametrize("variant,code,F", [
    ("decodeMinSum", "PEG", 300), ("decodeNormalizedMinSum", "802_3_H", 96), ("decodeOffsetMinSum", "802_3_H", 96),
    ("decodeBP", "PEG", 64), ("decodeDDBMP", "PEG", 200), ("decodeGDBF", "PEG", 300), ("decodeSMNGDBF", "PEG", 200),
    ("decodeMNGDBF", "802_3_H", 60), ("decodeStochasticNGDBF", "PEG", 100), ("decodeRSMNGDBF", "802_3_H", 60),
    ("NGDBFhw", "802_3_H", 200)])
def test_simulate_counters_match_oracle(variant, code, F):
    """simulate(seed, frames) on the GPU == the oracle decoding the oracle's own regeneration of the
    same Philox channel: every counter and histogram."""
    R, snr = cases.operating_point(variant, code)
    cfg = cases.cfg_for(variant, code=code)
    orc = Oracle(code)
    dec = capi.Decoder(capi.Code(code_path(code)), cfg)
    cws = load_codewords(os.path.join(os.path.dirname(code_path(code)), "data.enc"), 11) if code == "PEG" else None
    dec.set_codewords(cws)
    a = orc.simulate(cfg, snr, R, 99, 5000, F, cws)
    b = dec.simulate(snr, R, 99, 5000, F)
    if cfg.kind == abi.KIND_BP:
        for k in ("totalWords", "totalBits", "totalIterations", "uncodedErrors"):
            assert a.counters[k] == b.counters[k]
        assert abs(a.counters["errors"] - b.counters["errors"]) <= max(3, 0.02 * a.counters["errors"])
        return
    assert a.counters == b.counters
    assert np.array_equal(a.error_weight_hist, b.error_weight_hist)
    assert np.array_equal(a.iter_hist, b.iter_hist)
    assert np.array_equal(a.phase_hist, b.phase_hist)


def test_sharded_frame_ranges_add_up():
    """(e): the union of disjoint frame-id ranges equals one run over the whole range."""
    cfg = cases.cfg_for("decodeSMNGDBF")
    dec = capi.Decoder(capi.Code(code_path("PEG")), cfg)
    whole = dec.simulate(4.0, 0.5, 7, 0, 400).counters
    parts = [dec.simulate(4.0, 0.5, 7, lo, n).counters for lo, n in ((0, 100), (100, 37), (137, 263))]
    assert whole == {k: sum(p[k] for p in parts) for k in whole}


def test_stop_rule_polls():
    """Reference loop condition errors<200 || wordErrors<W (src/decodeMinSum.cpp:189), polled per launch."""
    cfg = cases.cfg_for("decodeMinSum", num_iterations=5)
    dec = capi.Decoder(capi.Code(code_path("PEG")), cfg)
    r = dec.simulate(1.5, 0.5, 3, 0, 10 ** 7, stop_errors=200, stop_word_errors=40, poll_frames=64).counters
    assert r["errors"] >= 200 and r["wordErrors"] >= 40
    assert r["totalWords"] < 10 ** 5 and r["totalWords"] % 64 == 0


@pytest.mark.gpu
@pytest.mark.parametrize("variant,code,T,snr", [("decodeRSMNGDBF", "PEG", 30, 4.0), ("decodeSMNGDBF", "802_3_H", 25, 4.0),
                                                 ("decodeStochasticNGDBF", "PEG", 12, 4.0), ("NGDBFhw", "802_3_H", 40, 4.3)])
def test_redecode_stats_matrix_matches_oracle(variant, code, T, snr):
    """SURVEY.md 8(f) N2.  ldpc_gpu_redecode_stats decodes every frame NR times from the same Philox channel samples
    with fresh decoder noise (re-decode r draws rows r*rows_per_decode ... of the frame's decoder stream).  The whole
    [frames][NR] matrix of error weights equals the double oracle's on the dumped samples and noise rows, and the
    counters equal the sum over all decodes.  (The oracle's re-decode semantics are pinned against the reference's
    redecodeStatistics main(): tests/test_oracle_vs_ref_main.py.)"""
    R = cases.CODES[code][0]
    NR, F = 3, 7
    over = dict(num_iterations=T)
    if variant == "decodeRSMNGDBF":
        over["maxphase"] = 1
    cfg = cases.cfg_for(variant, code=code, **over)
    dec = capi.Decoder(capi.Code(code_path(code)), cfg)
    got, cnt = dec.redecode_stats(snr, R, 99, 5, F, NR)
    rows = abi.noise_rows_needed(cfg) if cfg.kind == abi.KIND_GDBF else 0
    orc = Oracle(code)
    if cfg.kind == abi.KIND_NGDBF_HW:
        # NGDBFhw: the noise buffer of re-decode r is row r of the decoder stream; channel_dump returns row 0, so the oracle
        # checks column 0 and the other columns are checked for determinism and for differing from column 0 somewhere
        y, noise = dec.channel_dump(snr, R, 99, 5, F)
        a = orc.decode(cfg, snr, R, y, noise, 0, None)
        assert np.array_equal(got[:, 0], a.errors)
        again, _ = dec.redecode_stats(snr, R, 99, 5, F, NR)
        assert np.array_equal(got, again) and cnt["totalWords"] == F * NR
        return
    y, noise = dec.channel_dump(snr, R, 99, 5, F, noise_rows=rows * NR)
    want = np.zeros((F, NR), np.int32)
    for r in range(NR):
        a = orc.decode(cfg, snr, R, y, np.ascontiguousarray(noise[:, r * rows:(r + 1) * rows, :]), rows, None)
        want[:, r] = a.errors
    assert np.array_equal(got, want)
    assert cnt["totalWords"] == F * NR and cnt["errors"] == int(want.sum()) and cnt["wordErrors"] == int((want > 0).sum())
    if variant == "decodeRSMNGDBF":
        assert (want > 0).any() and (want == 0).any()          # both outcomes occur at this operating point
