"""The C++ host mains keep the reference binaries' entry points: positional command lines (usage text
compared with the reference's own main()), the usage-and-return-0 behaviour on a wrong argument count,
and the appended tab-separated result line (column layout compared with a real run of the reference's
main() on the harness's deterministic random() stream)."""
import os
import subprocess

import numpy as np
import pytest

import cases
from ldpcsimulation_b200 import abi, capi
from oracle.oracle_api import Reference, code_path

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
BIN = os.path.join(ROOT, "bin")

HOST_VARIANTS = ["decodeMinSum", "decodeOffsetMinSum", "decodeNormalizedMinSum", "decodeBP", "decodeDDBMP", "decodeGDBF",
                 "decodeMGDBF", "decodeSGDBF", "decodeStochasticNGDBF", "decodeMNGDBF", "decodeSMNGDBF", "decodeRSMNGDBF",
                 "decodeSMGDBF", "decodeSATGDBF", "decodeATGDBF", "NGDBFhw"]

needs_bin = pytest.mark.skipif(not os.path.exists(os.path.join(BIN, "ldpcsim")), reason="host mains not built")


def exe(variant):
    """bin/<variant> is a link to bin/ldpcsim; fall back to the `ldpcsim <variant>` form if links did not travel."""
    p = os.path.join(BIN, variant)
    return [p] if os.path.exists(p) else [os.path.join(BIN, "ldpcsim"), variant]


def _usage(text):
    line = [l for l in text.splitlines() if l.startswith("Usage:")][0]
    return line.split()[2:]            # drop "Usage:" and argv[0]


@needs_bin
@pytest.mark.ref
@pytest.mark.skipif(not Reference.available(), reason="oracle/_ref not built")
@pytest.mark.parametrize("variant", HOST_VARIANTS)
def test_usage_matches_reference_main(variant, tmp_path):
    out = tmp_path / "ref_stdout.txt"
    assert Reference(variant, "PEG").run_main([variant], 1, str(out)) == 0       # wrong argc: usage, return 0
    r = subprocess.run(exe(variant), capture_output=True, text=True)
    assert r.returncode == 0
    assert _usage(r.stdout) == _usage(out.read_text())


@needs_bin
def test_dispatch_by_first_argument_and_unknown_variant():
    r = subprocess.run([os.path.join(BIN, "ldpcsim"), "decodeBP"], capture_output=True, text=True)
    assert r.returncode == 0 and _usage(r.stdout) == ["alist", "R", "SNR", "T", "logfilename", "[codeword", "filename]"]
    r = subprocess.run([os.path.join(BIN, "ldpcsim"), "nonsense"], capture_output=True, text=True)
    assert r.returncode == 0 and "variants:" in r.stdout


@needs_bin
def test_bad_alist_is_reported(tmp_path):
    r = subprocess.run(exe("decodeMinSum") + [str(tmp_path / "missing.alist"), "0.5", "2.0", "5", str(tmp_path / "log")],
                       capture_output=True, text=True)
    assert r.returncode == 1 and "cannot open" in r.stderr


# args after `alist R SNR T` for a short run of each variant (PEG code)
ARGS = {
    "decodeMinSum": lambda log: [log],
    "decodeOffsetMinSum": lambda log: ["1.9375", "5", "0.125", log],
    "decodeNormalizedMinSum": lambda log: ["2.0", "6", "1.25", log],
    "decodeBP": lambda log: [log],
    "decodeDDBMP": lambda log: ["1.5", "4", log],
    "decodeGDBF": lambda log: ["-0.6", log],
    "decodeSMNGDBF": lambda log: ["-0.9", log, "0.975", "0.988", "1.0", "16", "2.5"],
    "decodeStochasticNGDBF": lambda log: ["-0.9", log, "0.9", "6", "1.0", "2.5"],
    "decodeRSMNGDBF": lambda log: ["-0.9", log, "0.975", "0.988", "1.0", "16", "2.5", "3"],
}


@needs_bin
@pytest.mark.gpu
@pytest.mark.parametrize("variant", sorted(ARGS))
def test_tsv_line_layout_and_values(variant, tmp_path):
    alist = code_path("PEG")
    log = str(tmp_path / "gpu.tsv")
    T = "6" if variant in ("decodeBP",) else "20"
    snr = "1.0"
    env = dict(os.environ, LDPC_SEED="77", LDPC_FRAMES="256", LDPC_POLL="256")
    cw = os.path.join(os.path.dirname(alist), "data.enc")
    r = subprocess.run(exe(variant) + [alist, "0.5", snr, T] + ARGS[variant](log) + [cw], capture_output=True, text=True, env=env)
    assert r.returncode == 0, r.stderr
    assert "Final result:" in r.stdout and "Using codewords from" in r.stdout
    mine = open(log).read().rstrip("\n").split("\t")
    # the same run through the C ABI directly
    kind, macros, _ = cases.VARIANTS[variant]
    names = [a for a in subprocess.run(exe(variant), capture_output=True, text=True).stdout.split()[2:]
             if a not in ("[codeword", "filename]")]
    vals = dict(zip(names, [alist, "0.5", snr, T] + ARGS[variant](log)))
    over = {k: float(vals[k]) for k in ("Ymax", "alpha", "delta", "theta", "noiseScale", "lambda") if k in vals}
    over.update({k: int(vals[k]) for k in ("Q", "NQ", "windowsize", "maxphase") if k in vals})
    cfg = abi.default_cfg(kind, flags=macros, num_iterations=int(T), **over)
    dec = capi.Decoder(capi.Code(alist), cfg)
    from oracle.oracle_api import load_codewords
    dec.set_codewords(load_codewords(cw))
    c = dec.simulate(float(snr), 0.5, 77, 0, 256).counters
    assert float(mine[1]) == pytest.approx(c["errors"] / c["totalBits"], rel=1e-5)
    assert float(mine[2]) == pytest.approx(c["totalIterations"] / c["totalWords"], rel=1e-5)
    assert float(mine[3]) == pytest.approx(c["wordErrors"] / c["totalWords"], rel=1e-5)
    assert mine[0] == "1" and mine[-1] == alist
    if not Reference.available(variant):
        return
    # column layout against a real run of the reference's own main()
    rlog = str(tmp_path / "ref.tsv")
    argv = [variant, alist, "0.5", snr, T] + [a if a != log else rlog for a in ARGS[variant](log)] + [cw]
    assert Reference(variant, "PEG").run_main(argv, 12345) == 0
    ref = open(rlog).read().rstrip("\n").split("\t")
    assert len(ref) == len(mine)
    stat_cols = {1, 2, 3}
    if kind == abi.KIND_GDBF:
        stat_cols |= {4, 5}
        if "outputSmoothing" in macros:
            base = 8 + sum(m in macros or (m == "addNoise" and "quantizeProbabilities" in macros)
                           for m in ("addNoise", "thresholdAdaptation", "weightSyndromes")) + (1 if ("quantizeSamples" in macros and "redecode" not in macros) else 0)
            stat_cols |= {base, base + 1}
    for i, (a, b) in enumerate(zip(ref, mine)):
        if i not in stat_cols:
            assert a == b, (i, ref, mine)          # SNR, T, parameters, alist: identical text


@needs_bin
@pytest.mark.gpu
def test_f16x2_host_main_writes_the_f64_line(tmp_path):
    """LDPC_PRECISION=f16x2 on decodeOffsetMinSum's exact-lattice macro set (the headline kernel behind the reference's CLI):
    the TSV line is the one LDPC_PRECISION=f64 writes, character for character."""
    alist = code_path("802_3_H")
    lines = {}
    for prec in ("f64", "f16x2"):
        log = str(tmp_path / (prec + ".tsv"))
        env = dict(os.environ, LDPC_SEED="5", LDPC_FRAMES="3000", LDPC_POLL="3000", LDPC_PRECISION=prec)
        r = subprocess.run(exe("decodeOffsetMinSum") + [alist, "0.8413", "3.6", "10", "1.9375", "5", "0.125", log], capture_output=True, text=True, env=env)
        assert r.returncode == 0, r.stderr
        if prec == "f16x2":
            assert "exact lattice" in r.stderr and "NOT" not in r.stderr
        lines[prec] = open(log).read()
    assert lines["f64"] == lines["f16x2"] and float(lines["f64"].split("\t")[3]) > 0      # word errors were part of the comparison


@needs_bin
@pytest.mark.gpu
def test_ngdbfhw_main(tmp_path):
    alist = code_path("802_3_H")
    log = str(tmp_path / "hw.tsv")
    r = subprocess.run(exe("NGDBFhw") + [alist, "4.5", "500", "1234", log], capture_output=True, text=True)
    assert r.returncode == 0, r.stderr
    cols = open(log).read().rstrip("\n").split("\t")
    # SNR errors wordErrors BER avgIt FER totalBits totalWords T theta0 noiseScale w Ymax NQ maxPhases seed (src/NGDBFhw.cpp:451-459)
    assert len(cols) == 16
    assert cols[0] == "4.5" and cols[6] == str(500 * 2048) and cols[7] == "500"
    assert cols[8:] == ["600", "-0.525", "0.95", "0.185", "1.625", "5", "1", "1234"]
    itdist = np.loadtxt(log + "_4.5_itdist.dat")
    assert itdist.shape == (600, 2) and itdist[0, 1] == 1.0 and np.all(np.diff(itdist[:, 1]) <= 0)
    if Reference.available("NGDBFhw"):
        rlog = str(tmp_path / "ref.tsv")
        assert Reference("NGDBFhw", "802_3_H").run_main(["NGDBFhw", alist, "4.5", "40", "1234", rlog], 99) == 0
        ref = open(rlog).read().rstrip("\n").split("\t")
        assert len(ref) == 16 and ref[8:] == cols[8:]


@needs_bin
@pytest.mark.ref
@pytest.mark.skipif(not Reference.available("redecodeStatistics"), reason="oracle/_ref not built")
def test_redecode_statistics_usage_matches_reference_main(tmp_path):
    out = tmp_path / "ref_stdout.txt"
    assert Reference.main("redecodeStatistics", ["redecodeStatistics"], 1, str(out)) == 0
    r = subprocess.run(exe("redecodeStatistics"), capture_output=True, text=True)
    assert r.returncode == 0 and _usage(r.stdout) == _usage(out.read_text())


@needs_bin
@pytest.mark.gpu
def test_redecode_statistics_rows(tmp_path):
    """bin/redecodeStatistics appends one row of NR tab-terminated error weights per frame (src/redecodeStatistics.cpp:
    392-395, 615-621), equal to ldpc_gpu_redecode_stats through the Python binding for the same seed."""
    log = tmp_path / "outcomes.txt"
    T, NR, NF, snr, R = 30, 5, 23, 4.0, 0.5
    env = dict(os.environ, LDPC_SEED="4711")
    argv = exe("redecodeStatistics") + [code_path("PEG"), str(R), str(snr), str(T), str(NR), str(NF), "-0.9", str(log),
                                        "0.975", "0.988", "1.0", "8", "2.5"]
    r = subprocess.run(argv, capture_output=True, text=True, env=env)
    assert r.returncode == 0, r.stderr
    text = log.read_text()
    assert all(line.endswith("\t") for line in text.splitlines())
    rows = np.array([[int(t) for t in line.split()] for line in text.splitlines()], np.int32)
    assert rows.shape == (NF, NR)
    cfg = abi.default_cfg(abi.KIND_GDBF, flags=["redecode", "addNoise", "thresholdAdaptation", "weightSyndromes", "outputSmoothing", "saturateSamples"],
                          num_iterations=T, theta=-0.9, noiseScale=0.975, alpha=1.0, windowsize=8, Ymax=2.5, maxphase=1, **{"lambda": 0.988})
    want, cnt = capi.Decoder(capi.Code(code_path("PEG")), cfg).redecode_stats(snr, R, 4711, 0, NF, NR)
    assert np.array_equal(rows, want)
    assert "Final result: %d bit errs in %d words" % (int(want.sum()), NF) in r.stdout
