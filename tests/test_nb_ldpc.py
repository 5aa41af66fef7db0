"""Non-binary GF(q) LDPC, min-max decoding (BASELINE.json configs[4]).  PARITY UNPINNED: the reference's
SystemC/NB-LDPC tree does not compile and holds no min-max decoder; the checker is the C restatement of the published
algorithm (oracle/ldpc_nb_oracle.c).  Fixtures: the reference's own GF(4) / GF(8) code files (codes/NB/, verbatim) and a
GF(16) code made by tools/make_nb_code.py (the reference ships none)."""
import os

import numpy as np
import pytest

from ldpcsimulation_b200 import capi
from oracle.oracle_api import NbOracle

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
CODES = {"gf4": ("codes/NB/q4.sp.9000.6000.4500.1", 1.0 / 3.0), "gf8": ("codes/NB/q8.sp.6000.4000.3000.1", 1.0 / 3.0),
         "gf16": ("codes/NB/gf16.reg.1536.768.alist", 0.5)}


def _path(k):
    return os.path.join(ROOT, CODES[k][0])


def _samples(orc, snr_db, R, F, seed, codeword_zero=True):
    sigma = np.sqrt(10 ** (-snr_db / 10) / R / 2)
    return 1.0 + sigma * np.random.default_rng(seed).standard_normal((F, orc.N * orc.m))


def test_nb_loader_and_gf_tables():
    for k, q in (("gf4", 4), ("gf8", 8), ("gf16", 16)):
        o = NbOracle(_path(k))
        assert o.q == q and o.m == q.bit_length() - 1
    c = capi.NbCode(_path("gf16"))
    assert (c.N, c.M, c.q, c.E) == (1536, 768, 16, 3072)


@pytest.mark.parametrize("k,snr", [("gf4", 2.0), ("gf16", 3.5)])
def test_nb_restatement_decodes(k, snr):
    """The restatement itself: clean frames stop at once; at a workable SNR it corrects the channel's symbol errors."""
    o = NbOracle(_path(k))
    clean = o.decode(5, np.ones((1, o.N * o.m)))
    assert clean.iters[0] == 0 and clean.counters["errors"] == 0
    y = _samples(o, snr, CODES[k][1], 2, 7)
    raw_sym_err = ((y.reshape(2, o.N, o.m) <= 0).any(axis=2)).sum()
    out = o.decode(30, y)
    assert raw_sym_err > 20 and out.counters["smoothingUsed"] < raw_sym_err / 4, (raw_sym_err, out.counters)


@pytest.mark.gpu
@pytest.mark.parametrize("k,snr,T,F", [("gf4", 1.6, 12, 3), ("gf8", 1.4, 10, 3), ("gf16", 3.0, 15, 12)])
def test_nb_kernel_equals_restatement(k, snr, T, F):
    o = NbOracle(_path(k))
    dec = capi.NbDecoder(capi.NbCode(_path(k)), T)
    y = _samples(o, snr, CODES[k][1], F, 11)
    a, b = o.decode(T, y), dec.decode(snr, CODES[k][1], y)
    assert np.array_equal(a.symbols, b.symbols) and np.array_equal(a.iters, b.iters) and a.counters == b.counters
    assert a.iters.max() > 0
    s1 = dec.simulate(snr, CODES[k][1], 5, 100, 2 * F).counters
    assert s1 == o.simulate(T, snr, CODES[k][1], 5, 100, 2 * F).counters
