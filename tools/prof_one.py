#!/usr/bin/env python3
"""One ldpc_gpu_simulate call for profiling.  Usage: prof_one.py <x2|h2rc|f32|f64> [frames] [T] [channel] [workload oms|nms]"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from ldpcsimulation_b200 import abi, capi  # noqa: E402

mode = sys.argv[1]
F = int(sys.argv[2]) if len(sys.argv) > 2 else 65536
T = int(sys.argv[3]) if len(sys.argv) > 3 else 10
ch = int(sys.argv[4]) if len(sys.argv) > 4 else 0
wl = sys.argv[5] if len(sys.argv) > 5 else "oms"
if mode == "h2rc":
    os.environ["LDPC_GPU_NO_X2"] = "1"
prec = {"x2": abi.PREC_F16X2, "h2rc": abi.PREC_F16X2, "f32": abi.PREC_F32, "f64": abi.PREC_F64}[mode]
kw = dict(flags=["quantizeSamples", "offsetMS"], Ymax=1.9375, Q=5, delta=0.125) if wl == "oms" else \
    dict(flags=["quantizeSamples", "normalizedMS"], Ymax=2.0, Q=6, alpha=1.25)
code = capi.Code(os.path.join(ROOT, "codes", "802_3", "802_3_H.alist"))
dec = capi.Decoder(code, abi.default_cfg(abi.KIND_MINSUM, num_iterations=T, precision=prec, channel_mode=ch, **kw))
for rep in range(3):
    r = dec.simulate(4.0, 0.8413, 1, rep * F, F)
print(mode, F, T, dec.last_timing(), r.counters, dec.stats())
