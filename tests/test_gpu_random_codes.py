"""Random small irregular parity-check matrices with awkward sizes (N not a multiple of 4 or 32, weight-1
columns, weight-2 rows, a dense row): exercises every tail / padding path of the generic kernels, which the
shipped codes (all multiples of 8, mostly regular) cannot.  fp64 must match the oracle bit for bit."""
import numpy as np
import pytest

import cases
from ldpcsimulation_b200 import abi, capi
from oracle.oracle_api import Oracle

pytestmark = pytest.mark.gpu


def random_alist(path, N, M, seed, dense_row=False):
    rng = np.random.default_rng(seed)
    H = np.zeros((M, N), np.uint8)
    for i in range(N):                                  # column weights 1..4
        w = int(rng.integers(1, 5))
        H[rng.choice(M, size=min(w, M), replace=False), i] = 1
    for j in range(M):                                  # every row at least weight 2
        while H[j].sum() < 2:
            H[j, rng.integers(0, N)] = 1
    if dense_row:
        H[0, rng.choice(N, size=min(N, 40), replace=False)] = 1
    cols = [np.flatnonzero(H[:, i]) + 1 for i in range(N)]
    rows = [np.flatnonzero(H[j]) + 1 for j in range(M)]
    dv, dc = max(map(len, cols)), max(map(len, rows))
    with open(path, "w") as f:
        f.write("%d %d\n%d %d\n" % (N, M, dv, dc))
        f.write(" ".join(str(len(c)) for c in cols) + "\n")
        f.write(" ".join(str(len(r)) for r in rows) + "\n")
        for c in cols:
            f.write(" ".join(map(str, list(c) + [0] * (dv - len(c)))) + "\n")
        for r in rows:
            f.write(" ".join(map(str, list(r) + [0] * (dc - len(r)))) + "\n")
    return dv, dc


VARIANTS = ["decodeMinSum", "decodeOffsetMinSum", "decodeNormalizedMinSum", "decodeBP", "decodeDDBMP", "decodeGDBF",
            "decodeSGDBF", "decodeMGDBF", "decodeSMNGDBF", "decodeStochasticNGDBF", "decodeRSMNGDBF", "decodeSeqATGDBF", "NGDBFhw"]


@pytest.mark.parametrize("N,M,seed,dense", [(37, 20, 1, False), (101, 47, 2, True), (258, 129, 3, False), (1, 1, 4, False), (63, 33, 5, True)])
def test_every_decoder_on_random_irregular_codes(tmp_path, N, M, seed, dense):
    path = str(tmp_path / "rnd.alist")
    if N == 1:
        open(path, "w").write("1 1\n1 1\n1\n1\n1\n1\n")          # the smallest legal alist
    else:
        random_alist(path, N, M, seed, dense)
    orc = Oracle(path)
    code = capi.Code(path)
    assert (code.N, code.M, code.E) == (orc.N, orc.M, orc.E)
    for variant in VARIANTS:
        cfg = cases.cfg_for(variant, num_iterations=7)
        if cfg.kind == abi.KIND_BP and orc.dc_max > 64:
            continue
        dec = capi.Decoder(code, cfg)
        F = 9
        y, noise, rows, cw = cases.make_inputs(orc.N, cfg, 3.0, 0.5, F, 100 + seed)
        a = orc.decode(cfg, 3.0, 0.5, y, noise, rows)
        b = dec.decode(3.0, 0.5, y, noise, rows)
        assert np.array_equal(a.bits, b.bits), variant
        assert np.array_equal(a.iters, b.iters) and np.array_equal(a.errors, b.errors) and np.array_equal(a.flags, b.flags), variant
        assert a.counters == b.counters, variant
        if cfg.kind in (abi.KIND_MINSUM, abi.KIND_DDBMP):
            assert np.array_equal(a.soft, b.soft), variant
        elif cfg.kind == abi.KIND_BP:
            np.testing.assert_allclose(a.soft, b.soft, rtol=1e-9, atol=1e-12)
        # throughput entry on the same code: counters equal the oracle's regeneration of the channel
        s0 = orc.simulate(cfg, 3.0, 0.5, 9, 50, 13)
        s1 = dec.simulate(3.0, 0.5, 9, 50, 13)
        if cfg.kind != abi.KIND_BP:
            assert s0.counters == s1.counters, variant
        # and the fp32 instantiation at least runs and agrees on the raw channel
        d32 = capi.Decoder(code, cases.cfg_for(variant, num_iterations=7, precision=abi.PREC_F32))
        s2 = d32.simulate(3.0, 0.5, 9, 50, 13)
        assert s2.counters["totalWords"] == 13 and abs(s2.counters["uncodedErrors"] - s1.counters["uncodedErrors"]) <= 1, variant
