"""One CPU worker of the reference baseline: decodes a bounded sample of frames with the reference's
own object code (oracle/_ref, kind "reference") or, where that is absent, the C restatement
(kind "port"), and prints one JSON line with its frames/s.  TEST / BENCH INFRASTRUCTURE ONLY.

The reference's only parallelism model is one independent OS process per operating point
(C_implementations/scripts/*.sh); bench.py starts one of these workers per host core and sums them.
Timing covers the frame loop only (alist load, allocation and sample generation excluded).
"""
import argparse
import json
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--variant", default="decodeNormalizedMinSum")
    ap.add_argument("--code", default="802_3_H")
    ap.add_argument("--snr", type=float, default=4.0)
    ap.add_argument("--rate", type=float, default=0.8413)
    ap.add_argument("--iters", type=int, default=10)
    ap.add_argument("--frames", type=int, default=256)
    ap.add_argument("--reps", type=int, default=1)
    ap.add_argument("--seed", type=int, default=1)
    ap.add_argument("--cfg", default="{}")
    ap.add_argument("--lib-suffix", default="", help="'_g': the build with the reference's own flags (-g, no -O)")
    a = ap.parse_args()

    import numpy as np
    import cases
    from oracle.oracle_api import Oracle, Reference

    over = json.loads(a.cfg)
    over["num_iterations"] = a.iters
    cfg = cases.cfg_for(a.variant, code=a.code, **over)
    if a.lib_suffix and not Reference.available(a.variant + a.lib_suffix):
        raise SystemExit("oracle/_ref/libref_%s%s.so is not built" % (a.variant, a.lib_suffix))
    if Reference.available(a.variant + a.lib_suffix):
        eng, kind = Reference(a.variant + a.lib_suffix, a.code), "reference"
    else:
        eng, kind = Oracle(a.code), "port"
    y, noise, rows, cw = cases.make_inputs(eng.N, cfg, a.snr, a.rate, a.frames, a.seed)
    eng.decode(cfg, a.snr, a.rate, y[:2], None if noise is None else noise[:2], rows)      # warm caches / page in
    t0 = time.perf_counter()
    for _ in range(a.reps):
        out = eng.decode(cfg, a.snr, a.rate, y, noise, rows)
    dt = time.perf_counter() - t0
    print(json.dumps({"kind": kind, "frames": a.frames * a.reps, "seconds": dt, "N": eng.N,
                      "frames_per_s": a.frames * a.reps / dt, "errors": out.counters["errors"]}))


if __name__ == "__main__":
    main()
