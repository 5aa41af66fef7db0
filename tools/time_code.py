#!/usr/bin/env python3
"""Kernel time of ldpc_gpu_simulate for any decoder / code.
Usage: python tools/time_code.py VARIANT CODE T FRAMES [f32|f64|f16x2] [snr_db] [reps]
e.g.   python tools/time_code.py decodeMinSum dvbs2 10 4096 f32 1.6"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import cases  # noqa: E402
from ldpcsimulation_b200 import abi, capi  # noqa: E402

CODE_FILES = {"PEG": "PEGReg504x1008/PEGReg504x1008.alist", "802_3_H": "802_3/802_3_H.alist", "802_3": "802_3/802_3.alist",
              "4000": "4000.2000.4.244/4000.2000.4.244.alist", "4376": "4376.282.4.9598/4376.282.4.9598.alist",
              "dvbs2": "dvbs2_1_2/dvbs2_1_2.alist"}


def code_path(name):
    return os.path.join(ROOT, "codes", CODE_FILES.get(name, name))


variant, code_name, T, F = sys.argv[1], sys.argv[2], int(sys.argv[3]), int(sys.argv[4])
prec = {"f64": abi.PREC_F64, "f32": abi.PREC_F32, "f16x2": abi.PREC_F16X2}[sys.argv[5] if len(sys.argv) > 5 else "f32"]
R, snr = cases.operating_point(variant, code_name) if code_name in cases.CODES else (0.5, 1.6)
if len(sys.argv) > 6:
    snr = float(sys.argv[6])
reps = int(sys.argv[7]) if len(sys.argv) > 7 else 3
code = capi.Code(code_path(code_name))
cfg = cases.cfg_for(variant, precision=prec, code=code_name, num_iterations=T)
dec = capi.Decoder(code, cfg)
dec.simulate(snr, R, 1, 0, F)
best = 1e9
for rep in range(reps):
    c = dec.simulate(snr, R, 1, (rep + 1) * F, F)["counters"]
    best = min(best, dec.last_timing()[0])
its = c["totalIterations"] / max(1, c["totalWords"])
print("%s %s T=%d %s snr=%.2f: %.3f ms / %d frames  %.3f Mframes/s  %.2f Gbit/s  avg it %.1f  FER %.2e  geometry %s" % (
    variant, code_name, T, sys.argv[5] if len(sys.argv) > 5 else "f32", snr, best, F, F / best / 1e3, F * code.N / best / 1e6,
    its, c["wordErrors"] / max(1, c["totalWords"]), dec.geometry()))
