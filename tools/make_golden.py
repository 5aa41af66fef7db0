#!/usr/bin/env python3
"""Generate tests/golden/*.npz: outputs of the REFERENCE'S OWN OBJECT CODE (oracle/_ref) on seeded inputs.

The reference ships no golden vectors, so these are produced here, in the
container where /root/reference exists, and committed.  Inputs are not stored: they are regenerated
from (seed, frame range) by the oracle's counter-based channel, and their SHA-256 is stored so that
any drift of that generator is caught rather than silently changing the test.
"""
import hashlib
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

import cases  # noqa: E402
from ldpcsimulation_b200 import abi  # noqa: E402
from oracle.oracle_api import Oracle, Reference, code_path, load_codewords  # noqa: E402

OUT = os.path.join(ROOT, "tests", "golden")
SEED, FRAME0, F = 20261018, 4096, 6


def sha(a):
    return hashlib.sha256(np.ascontiguousarray(a).tobytes()).hexdigest() if a is not None else ""


def golden_inputs(orc, cfg, variant, code):
    R, snr = cases.operating_point(variant, code)
    cws = load_codewords(os.path.join(os.path.dirname(code_path(code)), "data.enc"), 9) if code == "PEG" else None
    y, noise = orc.channel_dump(cfg, snr, R, SEED, FRAME0, F, cws)
    cw = None if cws is None else np.ascontiguousarray(cws[(FRAME0 + np.arange(F)) % len(cws)])
    return R, snr, y, noise, cw


def main():
    os.makedirs(OUT, exist_ok=True)
    for variant in cases.VARIANTS:
        for code in (["802_3_H"] if variant == "NGDBFhw" else ["PEG", "802_3_H"]):
            cfg = cases.cfg_for(variant, code=code)
            if cfg.kind == abi.KIND_BP and code == "802_3_H":
                continue
            orc, ref = Oracle(code), Reference(variant, code)
            R, snr, y, noise, cw = golden_inputs(orc, cfg, variant, code)
            rows = abi.noise_rows_needed(cfg)
            qp = np.arange(F, dtype=np.int32) * 37 % 600 if cfg.kind == abi.KIND_NGDBF_HW else None
            r = ref.decode(cfg, snr, R, y, noise, rows, cw, qpointer0=qp)
            keep_soft = cfg.kind in (abi.KIND_MINSUM, abi.KIND_BP, abi.KIND_DDBMP)
            np.savez_compressed(
                os.path.join(OUT, "%s__%s.npz" % (variant, code)),
                seed=SEED, frame0=FRAME0, F=F, snr=snr, R=R, y_sha256=sha(y), noise_sha256=sha(noise),
                bits=r.bits, iters=r.iters, errors=r.errors, flags=r.flags,
                soft=r.soft if keep_soft else np.zeros(0), counters=np.array([r.counters[k] for k in abi.Counters.SCALARS], np.int64),
                error_weight_hist=r.error_weight_hist, iter_hist=r.iter_hist, phase_hist=r.phase_hist,
                qpointer0=qp if qp is not None else np.zeros(0, np.int32))
            print(variant, code, r.iters.tolist(), r.errors.tolist())


if __name__ == "__main__":
    main()
