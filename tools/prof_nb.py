import sys
sys.path.insert(0, '.')
from ldpcsimulation_b200 import capi
code = capi.NbCode("codes/NB/gf16.reg.1536.768.alist"); dec = capi.NbDecoder(code, 15)
dec.simulate(3.5, 0.5, 1, 0, 296)
r = dec.simulate(3.5, 0.5, 1, 1000, 1480)
print(r.kernel_ms, r.counters)
