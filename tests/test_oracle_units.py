"""Oracle building blocks on the CPU: the counter-based channel (Philox known answers, Box-Muller
statistics), the three sample quantisers (a3) and their documented edge cases."""
import numpy as np

from oracle.oracle_api import Oracle


def test_philox4x32_10_known_answers():
    # Random123 kat_vectors
    kats = [((0, 0, 0, 0), (0, 0), (0x6627e8d5, 0xe169c58d, 0xbc57ac4c, 0x9b00dbd8)),
            ((0xffffffff,) * 4, (0xffffffff,) * 2, (0x408f276d, 0x41c83b0e, 0xa20bc7c6, 0x6d5451fd)),
            ((0x243f6a88, 0x85a308d3, 0x13198a2e, 0x03707344), (0xa4093822, 0x299f31d0), (0xd16cfe09, 0x94fdcceb, 0x5001e420, 0x24126ea1))]
    for ctr, key, want in kats:
        assert tuple(int(x) for x in Oracle.philox(ctr, key)) == want


def test_box_muller_is_standard_normal():
    n = np.concatenate([Oracle.normal4(42, f, b) for f in range(40) for b in range(500)]).astype(np.float64)
    assert abs(n.mean()) < 0.01 and abs(n.std() - 1) < 0.01
    assert abs(np.mean(n ** 3)) < 0.03 and abs(np.mean(n ** 4) - 3) < 0.06
    assert 4.0 < np.abs(n).max() < 6.7                      # radius uses all 32 bits: |n| <= sqrt(2*32*ln2) = 6.66
    # streams, rows, frames and seeds are independent coordinates
    a = Oracle.normal4(1, 2, 3, 0, 0)
    for other in (Oracle.normal4(2, 2, 3, 0, 0), Oracle.normal4(1, 3, 3, 0, 0), Oracle.normal4(1, 2, 4, 0, 0),
                  Oracle.normal4(1, 2, 3, 1, 0), Oracle.normal4(1, 2, 3, 0, 1)):
        assert not np.array_equal(a, other)
    assert np.array_equal(a, Oracle.normal4(1, 2, 3, 0, 0))


def test_ms_quantiser_edges():
    L = Oracle.lib()
    Ymax, Nq = 2.0, 8.0                                     # step = 4/7
    step = 2 * Ymax / (Nq - 1)
    assert L.oracle_quantize_ms(3.0, Ymax, Nq) == Ymax and L.oracle_quantize_ms(-3.0, Ymax, Nq) == -Ymax
    assert L.oracle_quantize_ms(0.0, Ymax, Nq) == step      # no zero level, sgn(0) = +1  (decodeMinSum.cpp:486-487)
    assert L.oracle_quantize_ms(-1e-9, Ymax, Nq) == -step
    assert L.oracle_quantize_ms(1.2, Ymax, Nq) == 2 * step  # floor, not round
    for x in np.linspace(-2.5, 2.5, 101):
        q = L.oracle_quantize_ms(float(x), Ymax, Nq)
        assert abs(q) <= Ymax and q != 0 and (q > 0) == (x >= 0)


def test_gdbf_quantiser_edges():
    L = Oracle.lib()
    Ymax, NQ = 2.0, 3                                       # qmax = 4, lmax = 1, step = 0.5
    assert L.oracle_quantize_gdbf(0.74, Ymax, NQ) == 0.5 and L.oracle_quantize_gdbf(0.76, Ymax, NQ) == 1.0
    assert L.oracle_quantize_gdbf(-0.76, Ymax, NQ) == -1.0
    assert L.oracle_quantize_gdbf(0.1, Ymax, NQ) == 0.0     # has a zero level; sgn(0) = -1 gives -0.0
    assert np.signbit(L.oracle_quantize_gdbf(0.0, Ymax, NQ))


def test_hw_pack_unpack_table():
    L = Oracle.lib()
    Ymax, w = 1.625, 0.185                                  # src/NGDBFhw.cpp:50-51
    lmax = Ymax / (2 * w)
    assert [L.oracle_hw_unpack(k) for k in (0, 1, 15)] == [1, 3, 31]
    assert [L.oracle_hw_unpack(16 | k) for k in (0, 1, 15)] == [-1, -3, -31]
    assert L.oracle_hw_unpack(L.oracle_hw_pack(2.0, Ymax, w)) == 15          # theta (:175)
    assert L.oracle_hw_pack(lmax, Ymax, w) == 15 and L.oracle_hw_pack(-lmax, Ymax, w) == 31
    assert L.oracle_hw_pack(0.0, Ymax, w) == 16                                # sgn(0) = -1: zero packs as negative
    for x in np.linspace(-lmax, lmax, 200):
        v = L.oracle_hw_unpack(L.oracle_hw_pack(float(x), Ymax, w))
        assert v % 2 != 0 and (v > 0) == (x > 0) and abs(v) <= 31


def test_quantizer_levels_encode_the_reference_quantiser():
    """abi.quantizer_levels (the LDPC_GPU_DT_Q8 encoding) against the oracle's restatement of quantize()
    (src/decodeMinSum.cpp:480-489, pinned to the reference object code in test_oracle_vs_ref.py): level * step, +-Ymax
    when saturated, never zero, sign of the sample (+ for +-0)."""
    import cases
    from ldpcsimulation_b200 import abi
    from oracle.oracle_api import Oracle
    for variant in ("decodeNormalizedMinSum", "decodeOffsetMinSum"):
        cfg = cases.cfg_for(variant, code="802_3_H", num_iterations=0)
        orc = Oracle("802_3_H")
        y, _, _, _ = cases.make_inputs(2048, cfg, 3.0, 0.8413, 6, 99)
        y[0, :8] = [0.0, -0.0, 1e-9, -1e-9, cfg.Ymax, -cfg.Ymax, cfg.Ymax * (1 + 1e-12), -3 * cfg.Ymax]
        want = orc.decode(cfg, 3.0, 0.8413, y).soft                      # T = 0: the conditioned samples
        k = abi.quantizer_levels(y, cfg.Ymax, cfg.Q)
        step = 2 * cfg.Ymax / (2.0 ** cfg.Q - 1)
        got = np.where(np.abs(k) == 32, np.sign(k) * cfg.Ymax, k.astype(np.float64) * step)
        assert np.array_equal(got, want) and np.abs(k).min() >= 1
