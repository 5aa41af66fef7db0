#!/usr/bin/env python3
"""Kernel time of ldpc_gpu_simulate on the exact-lattice offset min-sum workload (802.3an H, Ymax 1.9375, Q 5, delta 0.125):
packed exact kernel (variants), the labelled f16x2 kernel, fp32, fp64; exact / fast channel; T = 0 separates the front end.
Usage: python tools/time_x2.py [frames]"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from ldpcsimulation_b200 import abi, capi  # noqa: E402

F = int(sys.argv[1]) if len(sys.argv) > 1 else 1 << 19
code = capi.Code(os.path.join(ROOT, "codes", "802_3", "802_3_H.alist"))


def run(label, prec, env, T, channel, extra_flags=()):
    for k in ("LDPC_GPU_X2_VARIANT", "LDPC_GPU_NO_X2"):
        os.environ.pop(k, None)
    os.environ.update(env)
    cfg = abi.default_cfg(abi.KIND_MINSUM, flags=["quantizeSamples", "offsetMS"] + list(extra_flags), num_iterations=T, precision=prec,
                          Ymax=1.9375, Q=5, delta=0.125, channel_mode=channel)
    dec = capi.Decoder(code, cfg)
    dec.simulate(4.0, 0.8413, 1, 0, F)
    best = 1e9
    for rep in range(3):
        r = dec.simulate(4.0, 0.8413, 1, (rep + 1) * F, F)
        best = min(best, dec.last_timing()[0])
    x2, redo = dec.stats()
    print("%-26s T=%2d ch=%d  %8.3f ms  %6.2f Gbit/s  fer %.5f  x2=%d redo=%d  %s" % (
        label, T, channel, best, F * code.N / best / 1e6, r.counters["wordErrors"] / r.counters["totalWords"], x2, redo, dec.geometry()), flush=True)


for T in (0, 10):
    for ch in (0, 1):
        run("x2 reload (variant 1)", abi.PREC_F16X2, {"LDPC_GPU_X2_VARIANT": "1"}, T, ch)
        run("x2 3 CTAs/SM (variant 2)", abi.PREC_F16X2, {"LDPC_GPU_X2_VARIANT": "2"}, T, ch)
        if T:
            run("x2 certified stop", abi.PREC_F16X2, {"LDPC_GPU_X2_VARIANT": "1"}, T, ch, ["certStop"])
        run("h2rc labelled", abi.PREC_F16X2, {"LDPC_GPU_NO_X2": "1"}, T, ch)
        run("fp32 ms_rc", abi.PREC_F32, {}, T, ch)
run("fp64 ms_rc", abi.PREC_F64, {}, 10, 0)
