"""Shared parity cases: one entry per reference binary / macro set (Makefile:24-71), with the
operating points the reference's scripts record (scripts/*.sh) scaled to test size."""
import numpy as np

from ldpcsimulation_b200 import abi

# variant -> (kind, macros, cfg overrides)
VARIANTS = {
    "decodeMinSum":           (abi.KIND_MINSUM, [], dict(num_iterations=8)),
    "decodeOffsetMinSum":     (abi.KIND_MINSUM, ["quantizeSamples", "offsetMS"], dict(num_iterations=8, Ymax=1.9375, Q=5, delta=0.125)),
    "decodeNormalizedMinSum": (abi.KIND_MINSUM, ["quantizeSamples", "normalizedMS"], dict(num_iterations=8, Ymax=2.0, Q=6, alpha=1.25)),
    "decodeSaturatedMinSum":  (abi.KIND_MINSUM, ["saturateSamples"], dict(num_iterations=6, Ymax=1.5)),
    "decodeBP":               (abi.KIND_BP, [], dict(num_iterations=6)),
    "decodeDDBMP":            (abi.KIND_DDBMP, [], dict(num_iterations=12, Ymax=1.5, Q=4)),
    "decodeGDBF":             (abi.KIND_GDBF, [], dict(num_iterations=40, theta=-0.6)),
    "decodeMGDBF":            (abi.KIND_GDBF, ["modeswitching"], dict(num_iterations=40, theta=-0.6)),
    "decodeSGDBF":            (abi.KIND_GDBF, ["sequentialmode"], dict(num_iterations=40, theta=-0.6)),
    "decodeStochasticNGDBF":  (abi.KIND_GDBF, ["quantizeSamples", "quantizeProbabilities", "weightSyndromes", "saturateSamples"],
                               dict(num_iterations=40, theta=-0.9, noiseScale=0.9, NQ=6, alpha=1.0, Ymax=2.5)),
    "decodeMNGDBF":           (abi.KIND_GDBF, ["addNoise", "thresholdAdaptation", "weightSyndromes", "saturateSamples"],
                               dict(num_iterations=60, theta=-0.9, noiseScale=0.975, **{"lambda": 0.988}, alpha=1.0, Ymax=2.5)),
    "decodeSMNGDBF":          (abi.KIND_GDBF, ["addNoise", "thresholdAdaptation", "weightSyndromes", "outputSmoothing", "saturateSamples"],
                               dict(num_iterations=60, theta=-0.9, noiseScale=0.975, **{"lambda": 0.988}, alpha=0.8, windowsize=16, Ymax=2.5)),
    "decodeSMGDBF":           (abi.KIND_GDBF, ["outputSmoothing"], dict(num_iterations=40, theta=-0.6, windowsize=8)),
    "decodeSATGDBF":          (abi.KIND_GDBF, ["thresholdAdaptation", "outputSmoothing"], dict(num_iterations=40, theta=-0.6, **{"lambda": 0.99}, windowsize=8)),
    "decodeATGDBF":           (abi.KIND_GDBF, ["thresholdAdaptation"], dict(num_iterations=40, theta=-0.6, **{"lambda": 0.99})),
    "decodeUniformMNGDBF":    (abi.KIND_GDBF, ["addNoise", "uniformNoise", "thresholdAdaptation", "weightSyndromes", "saturateSamples"],
                               dict(num_iterations=60, theta=-0.9, noiseScale=0.975, **{"lambda": 0.988}, alpha=1.0, Ymax=2.5)),
    "decodeShapedMNGDBF":     (abi.KIND_GDBF, ["addNoise", "noiseShaping", "thresholdAdaptation", "weightSyndromes", "saturateSamples"],
                               dict(num_iterations=60, theta=-0.9, noiseScale=0.7, **{"lambda": 0.988}, alpha=1.0, Ymax=2.5)),
    "decodeSeqATGDBF":        (abi.KIND_GDBF, ["sequentialmode", "thresholdAdaptation"], dict(num_iterations=30, theta=-0.6, **{"lambda": 0.99})),
    "decodeRSMNGDBF":         (abi.KIND_GDBF, ["redecode", "addNoise", "thresholdAdaptation", "weightSyndromes", "outputSmoothing", "saturateSamples"],
                               dict(num_iterations=40, theta=-0.9, noiseScale=0.975, **{"lambda": 0.988}, alpha=1.0, windowsize=16, Ymax=2.5, maxphase=3)),
    "NGDBFhw":                (abi.KIND_NGDBF_HW, [], dict(num_iterations=80)),
}

# code -> (R, Eb/N0 for message passing, Eb/N0 for bit flipping): waterfall-ish points where some
# frames fail and some converge, so that both the early-stop and the run-to-T paths are exercised
CODES = {
    "PEG": (0.5, 2.4, 4.0),
    "802_3_H": (0.8413, 3.6, 4.8),
    "802_3": (0.8413, 3.6, 4.8),
    "4000": (0.5, 2.8, 4.5),
    "4376": (0.9356, 5.0, 6.0),
}


def operating_point(variant, code):
    R, snr_mp, snr_bf = CODES[code]
    kind = VARIANTS[variant][0]
    if variant == "decodeDDBMP":
        return R, snr_bf - 0.5
    return R, (snr_bf if kind in (abi.KIND_GDBF, abi.KIND_NGDBF_HW) else snr_mp)


def cfg_for(variant, precision=abi.PREC_F64, code=None, **over):
    kind, macros, kw = VARIANTS[variant]
    kw = dict(kw)
    if code in ("802_3_H", "802_3") and "weightSyndromes" in macros and "redecode" not in macros:
        kw.update(alpha=0.3, theta=-0.525)   # dv=6: w=alpha must shrink (cf. NGDBFhw w=0.185, src/NGDBFhw.cpp:50)
    kw.update(over)
    return abi.default_cfg(kind, flags=macros, precision=precision, **kw)


def sigma_of(snr_db, R):
    N0 = 10.0 ** (-snr_db / 10.0) / R
    return np.sqrt(N0 / 2.0)


def make_inputs(N, cfg, snr_db, R, F, seed, codewords=None):
    """Raw channel samples y = x(1+sigma n) (src/decodeMinSum.cpp:216) and the decoder-side raw RNG
    outputs in ldpc_gpu_batch layout, from a seeded numpy stream."""
    rng = np.random.default_rng(seed)
    cw = None
    x = np.ones((F, N))
    if codewords is not None:
        cw = np.ascontiguousarray(codewords[np.arange(F) % len(codewords)])
        x = 1.0 - 2.0 * cw
    y = x * (1.0 + sigma_of(snr_db, R) * rng.standard_normal((F, N)))
    noise, rows = None, 0
    if cfg.kind == abi.KIND_NGDBF_HW:
        noise = rng.standard_normal((F, abi.HW_QBUF))
    elif cfg.kind == abi.KIND_GDBF:
        rows = abi.noise_rows_needed(cfg)
        if rows:
            rps = abi.gdbf_rows_per_step(cfg.flags)
            noise = np.empty((F, rows, N))
            for rr in range(rows):
                which = rr % rps
                uniform = True
                if which == 0 and cfg.flags & abi.F_ADDNOISE:
                    uniform = bool(cfg.flags & abi.F_UNIFORMNOISE)
                if uniform:   # ranu() lattice, inc/rand.h:12-13
                    r31 = rng.integers(0, 2 ** 31, size=(F, N))
                    noise[:, rr, :] = (1.0 + r31) / (2.0 + float(0x7fffffff))
                else:
                    noise[:, rr, :] = rng.standard_normal((F, N))
    return y, noise, rows, cw
