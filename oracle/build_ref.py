#!/usr/bin/env python3
"""Build the oracle: the C restatement (always) and, when the reference tree is present,
the reference's own object code behind oracle/ref_harness.cpp (oracle/_ref/libref_*.so).

TEST INFRASTRUCTURE ONLY.  Reference sources are compiled where they lie; nothing is copied.
The reference's own Makefile is not run: each variant is one g++ line with the macro set of
/root/reference/C_implementations/Makefile:24-71 plus -Dmain=ref_main.
"""
import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
REF = os.environ.get("LDPC_REFERENCE", "/root/reference/C_implementations")
OUT = os.path.join(HERE, "_ref")
BUILD = os.path.join(HERE, "_build")

# name -> (reference TU, harness selector, reference macros)   [Makefile line]
VARIANTS = {
    "decodeMinSum":            ("decodeMinSum.cpp", "HARNESS_MS", []),                                   # :58
    "decodeOffsetMinSum":      ("decodeMinSum.cpp", "HARNESS_MS", ["quantizeSamples", "offsetMS"]),      # :61
    "decodeNormalizedMinSum":  ("decodeMinSum.cpp", "HARNESS_MS", ["quantizeSamples", "normalizedMS"]),  # :64
    "decodeSaturatedMinSum":   ("decodeMinSum.cpp", "HARNESS_MS", ["saturateSamples"]),                  # macro documented at decodeMinSum.cpp:29, no goal
    "decodeBP":                ("decodeBP.cpp", "HARNESS_BP", []),                                       # :67
    "decodeDDBMP":             ("decodeDDBMP.cpp", "HARNESS_DDBMP", []),                                 # :70
    "decodeGDBF":              ("decodeGDBF.cpp", "HARNESS_GDBF", []),                                   # plain parallel GDBF, no goal
    "decodeMGDBF":             ("decodeGDBF.cpp", "HARNESS_GDBF", ["modeswitching"]),                    # :24
    "decodeSGDBF":             ("decodeGDBF.cpp", "HARNESS_GDBF", ["sequentialmode"]),                   # :27
    "decodeStochasticNGDBF":   ("decodeGDBF.cpp", "HARNESS_GDBF", ["quantizeSamples", "quantizeProbabilities", "weightSyndromes", "saturateSamples"]),  # :30
    "decodeMNGDBF":            ("decodeGDBF.cpp", "HARNESS_GDBF", ["addNoise", "thresholdAdaptation", "weightSyndromes", "saturateSamples"]),           # :33
    "decodeSMNGDBF":           ("decodeGDBF.cpp", "HARNESS_GDBF", ["addNoise", "thresholdAdaptation", "weightSyndromes", "outputSmoothing", "saturateSamples"]),  # :36
    "decodeSMGDBF":            ("decodeGDBF.cpp", "HARNESS_GDBF", ["outputSmoothing"]),                  # :49
    "decodeSATGDBF":           ("decodeGDBF.cpp", "HARNESS_GDBF", ["thresholdAdaptation", "outputSmoothing"]),   # :52
    "decodeATGDBF":            ("decodeGDBF.cpp", "HARNESS_GDBF", ["thresholdAdaptation"]),              # :55
    "decodeUniformMNGDBF":     ("decodeGDBF.cpp", "HARNESS_GDBF", ["addNoise", "uniformNoise", "thresholdAdaptation", "weightSyndromes", "saturateSamples"]),  # scripts/ngdbf_example_uniform_*.sh
    "decodeShapedMNGDBF":      ("decodeGDBF.cpp", "HARNESS_GDBF", ["addNoise", "noiseShaping", "thresholdAdaptation", "weightSyndromes", "saturateSamples"]),
    "decodeSeqATGDBF":         ("decodeGDBF.cpp", "HARNESS_GDBF", ["sequentialmode", "thresholdAdaptation"]),    # exercises the running-minimum quirk
    "decodeRSMNGDBF":          ("RNGDBF.cpp", "HARNESS_GDBF", ["redecode", "addNoise", "thresholdAdaptation", "weightSyndromes", "outputSmoothing", "saturateSamples"]),  # :45
    "NGDBFhw":                 ("NGDBFhw.cpp", "HARNESS_HW", []),                                        # scripts/demo_NGDBFhw_802_3.sh:24
    # per-frame outcomes over NR re-decodes.  The Makefile's `redecodeStatistics` goal builds newstat.cpp
    # (GSL); src/redecodeStatistics.cpp is the GSL-free program of the same purpose and compiles on its own.  main() only.
    "redecodeStatistics":      ("redecodeStatistics.cpp", "HARNESS_MAINONLY", ["addNoise", "thresholdAdaptation", "weightSyndromes", "outputSmoothing", "saturateSamples"]),
}


BENCH_VARIANTS = ("decodeMinSum", "decodeOffsetMinSum", "decodeNormalizedMinSum", "decodeBP", "NGDBFhw", "decodeSMNGDBF")


def run(cmd):
    r = subprocess.run(cmd, capture_output=True, text=True)
    if r.returncode != 0:
        sys.stderr.write(" ".join(cmd) + "\n" + r.stdout + r.stderr)
        raise RuntimeError("build failed: " + cmd[-1])


def newer(target, sources):
    if not os.path.exists(target):
        return False
    t = os.path.getmtime(target)
    return all(os.path.exists(s) and os.path.getmtime(s) <= t for s in sources)


def build_restatement(force=False):
    os.makedirs(BUILD, exist_ok=True)
    out = os.path.join(BUILD, "libldpc_oracle.so")
    srcs = [os.path.join(HERE, "ldpc_oracle.c"), os.path.join(HERE, "ldpc_nb_oracle.c"), os.path.join(HERE, "ldpc_oracle.h"),
            os.path.join(HERE, "..", "include", "ldpc_gpu.h")]
    if force or not newer(out, srcs):
        run(["gcc", "-O2", "-g", "-std=c11", "-D_GNU_SOURCE", "-ffp-contract=off", "-fPIC", "-shared", "-Wall",
             srcs[0], srcs[1], "-lm", "-o", out])
    return out


def build_reference(force=False, opt="-O2", suffix="", only=None):
    """Returns the list of built libraries ([] when the reference tree is absent).
    opt="-O0", suffix="_g": the reference's own flags (`CFLAGS = -g -I./inc`, no -O, Makefile:5-6) for bench.py's cpu_baseline_refflags."""
    if not os.path.isdir(os.path.join(REF, "src")):
        return []
    os.makedirs(OUT, exist_ok=True)
    built = []
    harness = os.path.join(HERE, "ref_harness.cpp")
    common = [os.path.join(REF, "src", f) for f in ("alist.cpp", "r.cpp", "nrutil.cpp")]
    for name, (tu, sel, macros) in VARIANTS.items():
        if only is not None and name not in only:
            continue
        out = os.path.join(OUT, "libref_%s%s.so" % (name, suffix))
        src = os.path.join(REF, "src", tu)
        if force or not newer(out, [harness, src] + common):
            cmd = ["g++", opt, "-g", "-w", "-fPIC", "-shared", "-ffp-contract=off", "-I" + os.path.join(REF, "inc"),
                   "-Dmain=ref_main", "-D" + sel] + ["-D" + m for m in macros] + \
                  [src] + common + [harness, "-Wl,-Bsymbolic", "-lm", "-o", out]
            run(cmd)
        built.append(out)
    return built


if __name__ == "__main__":
    force = "--force" in sys.argv
    print(build_restatement(force))
    libs = build_reference(force)
    build_reference(force, opt="-O0", suffix="_g", only=BENCH_VARIANTS)
    print("%d reference variants built into %s" % (len(libs), OUT) if libs else "reference tree absent: oracle/_ref not rebuilt")
