"""LDPC_GPU_KIND_NGDBF_SC: the NGDBF decoder as the reference's SystemC model runs it.
PARITY UNPINNED: SystemC is absent, so the model itself cannot be run; the checker is the C restatement
(oracle/ldpc_oracle.c:sc_frame, citing SystemC/NGDBF/inc/nodes.h / decoder.h / ldpcsim.h line by line)."""
import os

import numpy as np
import pytest

import cases
from ldpcsimulation_b200 import abi, capi
from oracle.oracle_api import Oracle, code_path


def _transposed_alist(src, dst):
    """Rewrite an alist as the SystemC trees store it: the alist of H^T (header `M N`, SystemC/NGDBF/src/ldpcsim.cpp:107-110)."""
    rows = [list(map(int, ln.split())) for ln in open(src) if ln.strip()]
    N, M = rows[0]
    dv, dc = rows[1]
    cols, chks = rows[4:4 + N], rows[4 + N:4 + N + M]
    with open(dst, "w") as f:
        f.write("%d %d\n%d %d\n" % (M, N, dc, dv))
        f.write(" ".join(map(str, rows[3])) + "\n" + " ".join(map(str, rows[2])) + "\n")
        for r in chks:
            f.write(" ".join(map(str, r)) + "\n")
        for r in cols:
            f.write(" ".join(map(str, r)) + "\n")


def test_transposed_alist_loader(tmp_path):
    dst = str(tmp_path / "peg_t.alist")
    _transposed_alist(code_path("PEG"), dst)
    a, b = capi.Code(code_path("PEG")), capi.Code(dst, transposed=True)
    assert (a.N, a.M, a.E, a.dv_max, a.dc_max) == (b.N, b.M, b.E, b.dv_max, b.dc_max)
    assert np.array_equal(a.random_codewords(3, 4), b.random_codewords(3, 4))          # same H: same encoder
    t = capi.Code(dst)                                                                 # read as it stands it is a valid alist of H^T
    assert (t.N, t.M) == (a.M, a.N)


def test_sc_restatement_behaviour():
    """Sanity of the restatement itself: a clean frame stops in the first clock with zero flip steps; at a workable SNR most
    frames are decoded; the reported word of a frame that never stops is the smoothing vote."""
    orc = Oracle("PEG")
    cfg = abi.default_cfg(abi.KIND_NGDBF_SC, num_iterations=60)
    nl = abi.sc_noise_len(cfg, orc.N)
    rng = np.random.default_rng(1)
    clean = orc.decode(cfg, 8.0, 0.5, np.ones((2, orc.N)), rng.standard_normal((2, nl)), nl)
    assert np.all(clean.iters == 0) and np.all(clean.errors == 0) and np.all(clean.flags & 1)
    y = 1.0 + cases.sigma_of(6.0, 0.5) * rng.standard_normal((64, orc.N))
    out = orc.decode(cfg, 6.0, 0.5, y, rng.standard_normal((64, nl)), nl)
    assert (out.errors == 0).mean() > 0.8 and out.iters.max() <= 60
    bad = orc.decode(cfg, -3.0, 0.5, 1.0 + 1.5 * rng.standard_normal((4, orc.N)), rng.standard_normal((4, nl)), nl)
    assert np.all(bad.iters == 60) and np.all(bad.flags & 2)                          # ran to T: smoothed


@pytest.mark.gpu
@pytest.mark.parametrize("code_name,snr,T", [("PEG", 5.0, 80), ("802_3_H", 5.5, 50), ("4000", 5.0, 40)])
def test_sc_kernel_equals_restatement(code_name, snr, T):
    R = cases.CODES[code_name][0]
    cfg = abi.default_cfg(abi.KIND_NGDBF_SC, num_iterations=T, alpha=(0.95 if code_name != "802_3_H" else 0.6))
    orc = Oracle(code_name)
    code = capi.Code(code_path(code_name))
    dec = capi.Decoder(code, cfg)
    nl = abi.sc_noise_len(cfg, orc.N)
    rng = np.random.default_rng(5)
    cws = code.random_codewords(2, 5)
    F = 37
    cw = cws[np.arange(F) % 5]
    y = (1.0 - 2.0 * cw) * (1.0 + cases.sigma_of(snr, R) * rng.standard_normal((F, orc.N)))
    noise = rng.standard_normal((F, nl))
    a = orc.decode(cfg, snr, R, y, noise, nl, cw)
    b = dec.decode(snr, R, y, noise, nl, cw, want_soft=False)
    assert np.array_equal(a.bits, b.bits) and np.array_equal(a.iters, b.iters) and np.array_equal(a.errors, b.errors)
    assert np.array_equal(a.flags, b.flags) and a.counters == b.counters and np.array_equal(a.iter_hist, b.iter_hist)
    assert 0 < (a.iters < T).sum() and (a.errors == 0).sum() > 0
    # the throughput entry: Philox channel + Philox noise chain, re-derived by the restatement
    assert dec.simulate(snr, R, 11, 500, 200).counters == orc.simulate(cfg, snr, R, 11, 500, 200).counters
    y2, n2 = dec.channel_dump(snr, R, 11, 500, 3)
    y3, n3 = orc.channel_dump(cfg, snr, R, 11, 500, 3)
    assert np.array_equal(y2, y3) and np.array_equal(n2, n3)
