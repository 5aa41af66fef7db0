set -x
timeout 900 python -m pytest tests/test_nb_ldpc.py -x -q 2>&1 | tail -12 > gpurun_out/r2aj_pytest_nb.log
python - > gpurun_out/r2aj_nb_time.log 2>&1 <<'PY'
import os, sys
sys.path.insert(0, '.')
from ldpcsimulation_b200 import capi
for path, R, snr, T in (("codes/NB/gf16.reg.1536.768.alist", 0.5, 3.5, 15), ("codes/NB/q4.sp.9000.6000.4500.1", 1/3, 2.0, 15), ("codes/NB/q8.sp.6000.4000.3000.1", 1/3, 2.0, 15)):
    code = capi.NbCode(path); dec = capi.NbDecoder(code, T)
    dec.simulate(snr, R, 1, 0, 296)
    for F in (2960, 8880):
        r = dec.simulate(snr, R, 1, 1000, F)
        c = r.counters
        print(path, "F", F, "ms %.2f" % r.kernel_ms, "Gbit/s %.3f" % (F * code.N * code.m / r.kernel_ms / 1e6), "avg it %.2f" % (c["totalIterations"] / c["totalWords"]), "FER %.4f" % (c["wordErrors"] / c["totalWords"]), "BER %.2e" % (c["errors"] / c["totalBits"]), flush=True)
PY
