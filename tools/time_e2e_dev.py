#!/usr/bin/env python3
"""Kernel time of ldpc_gpu_decode_batch with DEVICE buffers (bit-packed levels in, decisions + iteration counts out) next to the
HOST-buffer end-to-end time at several batch sizes: what the two-slot pipeline loses.  Usage: time_e2e_dev.py"""
import os
import sys
import time

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from ldpcsimulation_b200 import abi, capi  # noqa: E402

code = capi.Code(os.path.join(ROOT, "codes", "802_3", "802_3_H.alist"))
N = code.N
dec = capi.Decoder(code, abi.default_cfg(abi.KIND_MINSUM, flags=["quantizeSamples", "offsetMS"], num_iterations=10, precision=abi.PREC_F16X2,
                                         Ymax=1.9375, Q=5, delta=0.125))
F0 = 1 << 14
y = 1.0 + 0.4705 * np.random.default_rng(1).standard_normal((F0, N))
lev = torch.from_numpy(abi.quantizer_levels_packed(y, 1.9375, 5))
for F in (1 << 17, 1 << 19, 1 << 20):
    rep = F // F0
    yh = torch.empty((F, lev.shape[1]), dtype=torch.uint8, pin_memory=True); yh.copy_(lev.repeat(rep, 1))
    bits = torch.empty((F, N // 8), dtype=torch.uint8, pin_memory=True)
    iters = torch.empty((F,), dtype=torch.int32, pin_memory=True)
    yd, bd, itd = yh.cuda(), torch.empty((F, N // 8), dtype=torch.uint8, device="cuda"), torch.empty((F,), dtype=torch.int32, device="cuda")
    for mem, yy, bb, ii, name in ((abi.MEM_DEVICE, yd, bd, itd, "device"), (abi.MEM_HOST, yh, bits, iters, "host  ")):
        b = abi.Batch(); b.n_frames, b.mem, b.y_dtype = F, mem, abi.DT_QP
        b.y, b.out_bits, b.out_iters = yy.data_ptr(), bb.data_ptr(), ii.data_ptr()
        for _ in range(2):
            dec.decode_raw(4.0, 0.8413, b)
        torch.cuda.synchronize(); t = time.perf_counter()
        for _ in range(5):
            dec.decode_raw(4.0, 0.8413, b)
        torch.cuda.synchronize(); w = time.perf_counter() - t
        print("F %8d %s buffers: %6.2f Gbit/s" % (F, name, 5 * F * N / w / 1e9), flush=True)
