#!/usr/bin/env python3
"""Tiny pass over every kernel family (for compute-sanitizer): a few frames each, results checked
against the oracle where that is cheap."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np  # noqa: E402
import cases  # noqa: E402
from ldpcsimulation_b200 import abi, capi  # noqa: E402
from oracle.oracle_api import Oracle, code_path  # noqa: E402

JOBS = [("decodeNormalizedMinSum", "802_3_H", abi.PREC_F32), ("decodeNormalizedMinSum", "802_3_H", abi.PREC_F64),
        ("decodeNormalizedMinSum", "802_3_H", abi.PREC_F16X2), ("decodeOffsetMinSum", "802_3", abi.PREC_F32),
        ("decodeMinSum", "PEG", abi.PREC_F32), ("decodeMinSum", "4376", abi.PREC_F64), ("decodeBP", "802_3_H", abi.PREC_F32),
        ("decodeBP", "PEG", abi.PREC_F64), ("decodeDDBMP", "PEG", abi.PREC_F64), ("decodeSMNGDBF", "PEG", abi.PREC_F64),
        ("decodeRSMNGDBF", "802_3_H", abi.PREC_F32), ("decodeSGDBF", "PEG", abi.PREC_F64), ("decodeStochasticNGDBF", "PEG", abi.PREC_F64),
        ("NGDBFhw", "802_3_H", abi.PREC_F64), ("decodeMinSum", "dvbs2", abi.PREC_F32)]
for variant, code, prec in JOBS:
    R, snr = (0.5, 1.6) if code == "dvbs2" else cases.operating_point(variant, code)
    T = 4 if code == "dvbs2" else None
    cfg = cases.cfg_for(variant, code=code, precision=prec, **({"num_iterations": T} if T else {}))
    dec = capi.Decoder(capi.Code(code_path(code)), cfg)
    F = 5 if code != "dvbs2" else 9
    r = dec.simulate(snr, R, 3, 0, F)
    y, noise = dec.channel_dump(snr, R, 3, 0, 3)
    rows = abi.noise_rows_needed(cfg)
    out = dec.decode(snr, R, y, noise, rows)
    ok = ""
    if prec == abi.PREC_F64 and code != "dvbs2":
        cfg64 = cases.cfg_for(variant, code=code)
        ref = Oracle(code).decode(cfg64, snr, R, y, noise, rows)
        assert np.array_equal(ref.bits, out.bits) and ref.counters == out.counters, (variant, code)
        ok = "== oracle"
    print("%-24s %-8s prec %d  words %d errors %d %s" % (variant, code, prec, r.counters["totalWords"], r.counters["errors"], ok), flush=True)
print("sanitize smoke done")
