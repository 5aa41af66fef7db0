#!/usr/bin/env python3
"""e2e throughput of ldpc_gpu_decode_batch (HOST buffers, one-byte quantiser levels) vs chunk size.  Usage: time_e2e.py [frames]"""
import os
import sys
import time

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from ldpcsimulation_b200 import abi, capi  # noqa: E402

F = int(sys.argv[1]) if len(sys.argv) > 1 else 1 << 17
code = capi.Code(os.path.join(ROOT, "codes", "802_3", "802_3_H.alist"))
N = code.N
y = 1.0 + 0.5 * torch.randn((F, N))
lev = torch.from_numpy(abi.quantizer_levels(y.numpy(), 1.9375, 5))
yh = torch.empty((F, N), dtype=torch.int8, pin_memory=True); yh.copy_(lev)
bits = torch.empty((F, N // 8), dtype=torch.uint8, pin_memory=True)
iters = torch.empty((F,), dtype=torch.int32, pin_memory=True)
for prec, name in ((abi.PREC_F16X2, "x2"), (abi.PREC_F32, "f32")):
    for mb in (8, 16, 32, 64, 128):
        os.environ["LDPC_GPU_CHUNK_MB"] = str(mb)
        dec = capi.Decoder(code, abi.default_cfg(abi.KIND_MINSUM, flags=["quantizeSamples", "offsetMS"], num_iterations=10, precision=prec,
                                                 Ymax=1.9375, Q=5, delta=0.125))
        b = abi.Batch(); b.n_frames, b.mem, b.y_dtype = F, abi.MEM_HOST, abi.DT_Q8
        b.y, b.out_bits, b.out_iters = yh.data_ptr(), bits.data_ptr(), iters.data_ptr()
        for _ in range(2):
            dec.decode_raw(4.0, 0.8413, b)
        t = time.perf_counter(); kms = 0
        for _ in range(5):
            dec.decode_raw(4.0, 0.8413, b); kms += dec.last_timing()[0]
        w = time.perf_counter() - t
        print("%s chunk %3d MB: e2e %6.2f Gbit/s   kernel-only %6.2f Gbit/s  launches %d" % (name, mb, 5 * F * N / w / 1e9, 5 * F * N / kms / 1e6, dec.last_timing()[1]), flush=True)
