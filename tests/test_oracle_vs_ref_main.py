"""Whole-program pin: the reference's UNMODIFIED main() (its inline channel / conditioning code, frame
loop, stop rule, accounting and TSV line) run on the harness's deterministic random() stream, replayed
through the oracle frame by frame with the same stream.  What the reference wrote to its log must be
what the oracle's counters give.

rann() is `cos(2*pi*ranf()) * sqrt(-2*log(1-ranf()))` (inc/rand.h:19-20); the order in which the two
ranf() calls are evaluated is unspecified in C++ and fixed by the compiler.  With the g++ used to build
oracle/_ref the LEFT operand's ranf() (the cosine's) is drawn first; the test asserts that (and would fail
loudly, not silently, under a compiler that chooses the other order)."""
import math
import os

import numpy as np
import pytest

import cases
from ldpcsimulation_b200 import abi
from oracle.oracle_api import Oracle, Reference, code_path, load_codewords

pytestmark = [pytest.mark.ref, pytest.mark.skipif(not Reference.available(), reason="oracle/_ref not built")]

MASK = (1 << 64) - 1


class Shim:
    """The xorshift64 stream of oracle/ref_harness.cpp's random()."""

    def __init__(self, seed):
        self.s = seed if seed else 88172645463325252

    def random(self):
        s = self.s
        s ^= (s << 13) & MASK
        s ^= s >> 7
        s ^= (s << 17) & MASK
        self.s = s
        return (s >> 20) & 0x7fffffff

    def ranf(self):
        return self.random() / (1.0 + float(0x7fffffff))

    def rann(self):
        a = self.ranf()            # left operand (the cosine's) first, see module docstring
        b = self.ranf()
        return math.cos(2.0 * 3.141592654 * a) * math.sqrt(-2.0 * math.log(1.0 - b))

    def rann_rl(self):
        b = self.ranf()
        a = self.ranf()
        return math.cos(2.0 * 3.141592654 * a) * math.sqrt(-2.0 * math.log(1.0 - b))

    def copy(self):
        c = Shim(1)
        c.s = self.s
        return c


def replay(variant, code, argv_tail, cfg, R, snr, order, stream_seed, codewords, min_word_errors, fixed_frames=None):
    orc = Oracle(code)
    N = orc.N
    sh = Shim(stream_seed)
    draw = (lambda g: g.rann()) if order == "lr" else (lambda g: g.rann_rl())
    sigma = cases.sigma_of(snr, R)
    tot = dict(errors=0, wordErrors=0, totalWords=0, totalBits=0, totalIterations=0, smoothingUsed=0, uncodedErrors=0)
    rows_per_it = abi.gdbf_rows_per_step(cfg.flags) if cfg.kind == abi.KIND_GDBF else 0
    qp = 0
    f = 0
    while True:
        if fixed_frames is not None:
            if tot["totalWords"] >= fixed_frames:
                break
        elif not (tot["errors"] < 200 or tot["wordErrors"] < min_word_errors):
            break
        cw = None if codewords is None else codewords[f % len(codewords)][None, :]
        x = np.ones(N) if cw is None else 1.0 - 2.0 * cw[0]
        y = np.array([x[i] * (1.0 + sigma * draw(sh)) for i in range(N)])[None, :]
        noise, rows, qarg = None, 0, None
        if cfg.kind == abi.KIND_NGDBF_HW:
            noise = np.array([draw(sh) for _ in range(abi.HW_QBUF)])[None, :]
            qarg = np.array([qp], np.int32)
        elif rows_per_it:
            look = sh.copy()                                   # rows the decoder MAY consume; only executed ones advance the stream
            rows = abi.noise_rows_needed(cfg)
            noise = np.array([[draw(look) for _ in range(N)] for _ in range(rows)])[None, :, :]
        out = orc.decode(cfg, snr, R, y, noise, rows, cw, qpointer0=qarg)
        it = int(out.iters[0])
        if rows_per_it:
            for _ in range(it * N * rows_per_it):
                draw(sh)
        if cfg.kind == abi.KIND_NGDBF_HW:
            qp = (qp + it) % (abi.HW_QBUF - N)
        for k in tot:
            tot[k] += out.counters[k]
        f += 1
    return tot


def _fmt(v):
    return "%g" % v


CASES = [
    ("decodeMinSum", "PEG", 1.2, 8, lambda log: [log], True, 40),
    ("decodeNormalizedMinSum", "PEG", 1.4, 6, lambda log: ["2.0", "6", "1.25", log], False, 40),
    ("decodeOffsetMinSum", "PEG", 1.4, 6, lambda log: ["1.9375", "5", "0.125", log], True, 40),
    ("decodeDDBMP", "PEG", 2.5, 8, lambda log: ["1.5", "4", log], False, 40),
    ("decodeGDBF", "PEG", 2.5, 15, lambda log: ["-0.6", log], True, 20),
    ("decodeSMNGDBF", "PEG", 2.5, 12, lambda log: ["-0.9", log, "0.975", "0.988", "1.0", "4", "2.5"], False, 20),
    ("decodeBP", "PEG", 1.0, 5, lambda log: [log], True, 20),
    ("decodeRSMNGDBF", "PEG", 3.0, 10, lambda log: ["-0.9", log, "0.975", "0.988", "1.0", "4", "2.5", "3"], True, 20),
]


@pytest.mark.parametrize("variant,code,snr,T,tail,use_cw,minw", CASES)
def test_reference_main_totals(variant, code, snr, T, tail, use_cw, minw, tmp_path):
    R = 0.5
    alist = code_path(code)
    log = str(tmp_path / "ref.tsv")
    cwfile = os.path.join(os.path.dirname(alist), "data.enc")
    argv = [variant, alist, str(R), str(snr), str(T)] + tail(log) + ([cwfile] if use_cw else [])
    assert Reference(variant, code).run_main(argv, 4242) == 0
    ref = open(log).read().rstrip("\n").split("\t")
    kind, macros, _ = cases.VARIANTS[variant]
    names = ["alist", "R", "SNR", "T"]
    over = {}
    t = tail(log)
    if variant in ("decodeNormalizedMinSum",):
        over = dict(Ymax=float(t[0]), Q=int(t[1]), alpha=float(t[2]))
    elif variant == "decodeOffsetMinSum":
        over = dict(Ymax=float(t[0]), Q=int(t[1]), delta=float(t[2]))
    elif variant == "decodeDDBMP":
        over = dict(Ymax=float(t[0]), Q=int(t[1]))
    elif variant == "decodeGDBF":
        over = dict(theta=float(t[0]))
    elif variant == "decodeSMNGDBF":
        over = dict(theta=float(t[0]), noiseScale=float(t[2]), alpha=float(t[4]), windowsize=int(t[5]), Ymax=float(t[6]))
        over["lambda"] = float(t[3])
    elif variant == "decodeRSMNGDBF":
        over = dict(theta=float(t[0]), noiseScale=float(t[2]), alpha=float(t[4]), windowsize=int(t[5]), Ymax=float(t[6]), maxphase=int(t[7]))
        over["lambda"] = float(t[3])
    cfg = abi.default_cfg(kind, flags=macros, num_iterations=T, **over)
    cws = load_codewords(cwfile) if use_cw else None
    tot = replay(variant, code, t, cfg, R, snr, "lr", 4242, cws, minw)
    ber, avgit, fer = tot["errors"] / tot["totalBits"], tot["totalIterations"] / tot["totalWords"], tot["wordErrors"] / tot["totalWords"]
    assert ref[1:4] == [_fmt(ber), _fmt(avgit), _fmt(fer)], (ref, tot)
    if kind == abi.KIND_GDBF:
        assert ref[4:6] == [str(tot["totalBits"]), str(tot["totalWords"])]
        if "outputSmoothing" in macros:
            i = ref.index(str(cfg.windowsize), 8) - 2
            assert ref[i] == str(tot["smoothingUsed"])


def test_operand_order_is_detected():
    """The other evaluation order does NOT reproduce the reference: the pin is sensitive to it."""
    variant, code, snr, T = "decodeMinSum", "PEG", 1.2, 8
    import tempfile
    with tempfile.TemporaryDirectory() as d:
        log = os.path.join(d, "r.tsv")
        assert Reference(variant, code).run_main([variant, code_path(code), "0.5", str(snr), str(T), log], 7) == 0
        ref = open(log).read().split("\t")
    cfg = abi.default_cfg(abi.KIND_MINSUM, num_iterations=T)
    good = replay(variant, code, [], cfg, 0.5, snr, "lr", 7, None, 40)
    bad = replay(variant, code, [], cfg, 0.5, snr, "rl", 7, None, 40)
    assert ref[1] == _fmt(good["errors"] / good["totalBits"])
    assert good != bad and ref[1] != _fmt(bad["errors"] / bad["totalBits"])


def test_ngdbfhw_main_totals(tmp_path):
    """NGDBFhw: fixed frame count, raw error counts in the TSV, noise-window pointer carried across frames."""
    alist = code_path("802_3_H")
    log = str(tmp_path / "hw.tsv")
    assert Reference("NGDBFhw", "802_3_H").run_main(["NGDBFhw", alist, "4.0", "25", "1234", log], 99) == 0
    ref = open(log).read().rstrip("\n").split("\t")
    cfg = cases.cfg_for("NGDBFhw", num_iterations=600)
    tot = replay("NGDBFhw", "802_3_H", [], cfg, 0.8413, 4.0, "lr", 99, None, 0, fixed_frames=25)
    assert ref[1] == str(tot["errors"]) and ref[2] == str(tot["wordErrors"])
    assert ref[4] == _fmt(tot["totalIterations"] / tot["totalWords"])
    assert ref[6:8] == [str(tot["totalBits"]), str(tot["totalWords"])]


def test_redecode_statistics_outcome_matrix(tmp_path):
    """SURVEY.md 8(f) N2.  src/redecodeStatistics.cpp decodes every received frame NR times from the same channel
    samples with fresh perturbation noise and logs one row of NR error weights per frame.  Replayed on the same
    random() stream through the oracle: one RSMNGDBF decode with maxphase = 1 per re-decode (the file's symNodeUpdates
    uses RNGDBF's weight alpha*Ymax/dv, :543-546), noise rows consumed only for the iterations that ran."""
    code, R, snr, T, NR, NF = "PEG", 0.5, 4.2, 40, 4, 9
    log = str(tmp_path / "outcomes.txt")
    argv = ["redecodeStatistics", code_path(code), str(R), str(snr), str(T), str(NR), str(NF), "-0.9", log, "0.975", "0.988", "1.0", "4", "2.5"]
    assert Reference.main("redecodeStatistics", argv, 31337) == 0
    ref_rows = [[int(t) for t in line.split()] for line in open(log).read().strip().splitlines()]
    assert len(ref_rows) == NF and all(len(r) == NR for r in ref_rows)
    cfg = abi.default_cfg(abi.KIND_GDBF, flags=["redecode", "addNoise", "thresholdAdaptation", "weightSyndromes", "outputSmoothing", "saturateSamples"],
                          num_iterations=T, theta=-0.9, noiseScale=0.975, alpha=1.0, windowsize=4, Ymax=2.5, maxphase=1, **{"lambda": 0.988})
    orc = Oracle(code)
    N = orc.N
    sh = Shim(31337)
    sigma = cases.sigma_of(snr, R)
    mine = []
    for f in range(NF):
        y = np.array([1.0 + sigma * sh.rann() for _ in range(N)])[None, :]
        row = []
        for r in range(NR):
            look = sh.copy()
            noise = np.array([[look.rann() for _ in range(N)] for _ in range(T)])[None, :, :]
            out = orc.decode(cfg, snr, R, y, noise, T, None)
            for _ in range(int(out.iters[0]) * N):
                sh.rann()
            row.append(int(out.errors[0]))
        mine.append(row)
    assert mine == ref_rows
    assert any(any(r) for r in ref_rows) and any(not all(r) for r in ref_rows)       # both outcomes occur
