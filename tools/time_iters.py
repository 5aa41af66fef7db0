#!/usr/bin/env python3
"""Kernel time of ldpc_gpu_simulate vs iteration count: separates the channel front end / epilogue
(T=0) from the per-iteration cost.  Usage: python tools/time_iters.py [frames] [precision]"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from ldpcsimulation_b200 import abi, capi  # noqa: E402

F = int(sys.argv[1]) if len(sys.argv) > 1 else 1 << 18
prec = {"f64": abi.PREC_F64, "f32": abi.PREC_F32, "f16x2": abi.PREC_F16X2}[sys.argv[2] if len(sys.argv) > 2 else "f32"]
code = capi.Code(os.path.join(ROOT, "codes", "802_3", "802_3_H.alist"))
for T in (0, 1, 2, 10, 20):
    cfg = abi.default_cfg(abi.KIND_MINSUM, flags=["quantizeSamples", "normalizedMS"], num_iterations=T, precision=prec,
                          Ymax=2.0, Q=6, alpha=1.25)
    dec = capi.Decoder(code, cfg)
    dec.simulate(4.0, 0.8413, 1, 0, F)
    best = 1e9
    for rep in range(3):
        dec.simulate(4.0, 0.8413, 1, (rep + 1) * F, F)
        best = min(best, dec.last_timing()[0])
    print("T=%2d  %8.3f ms  %7.2f Mframes/s  %6.2f Gbit/s   geometry %s" % (T, best, F / best / 1e3, F * code.N / best / 1e6, dec.geometry()))
