"""Statistical agreement (north-star: BER/FER within Monte-Carlo 95 % confidence intervals).

The fp64 instantiation is bit-exact against the reference's object code (test_gpu_parity / test_golden),
so it stands in for the reference at frame counts the CPU cannot reach; the fp32 throughput
instantiation must then land inside the 95 % interval around it, on the same Philox channel."""
import math

import pytest

import cases
from ldpcsimulation_b200 import abi, capi
from oracle.oracle_api import code_path

pytestmark = pytest.mark.gpu


def _wilson(k, n, z=1.96):
    p = k / n
    d = 1 + z * z / n
    c = (p + z * z / (2 * n)) / d
    h = z * math.sqrt(p * (1 - p) / n + z * z / (4 * n * n)) / d
    return c - h, c + h


@pytest.mark.parametrize("variant,code,snr,F", [
    ("decodeNormalizedMinSum", "802_3_H", 3.6, 400000),
    ("decodeOffsetMinSum", "802_3_H", 3.8, 400000),
    ("decodeMinSum", "PEG", 2.4, 300000),
    ("decodeBP", "PEG", 2.0, 60000),
])
def test_fp32_fer_inside_fp64_confidence_interval(variant, code, snr, F):
    R = cases.CODES[code][0]
    code_h = capi.Code(code_path(code))
    r64 = capi.Decoder(code_h, cases.cfg_for(variant, code=code, num_iterations=10)).simulate(snr, R, 2026, 0, F).counters
    # a different frame range: independent noise, same distribution
    r32 = capi.Decoder(code_h, cases.cfg_for(variant, code=code, num_iterations=10, precision=abi.PREC_F32)).simulate(snr, R, 2026, F, F).counters
    assert r64["wordErrors"] >= 100, "operating point too clean for a meaningful interval"
    lo, hi = _wilson(r64["wordErrors"], F)
    # two independent estimates: widen by sqrt(2)
    mid, half = (lo + hi) / 2, (hi - lo) / 2 * math.sqrt(2)
    fer32 = r32["wordErrors"] / F
    assert mid - half <= fer32 <= mid + half, (r64, r32)
    ber64, ber32 = r64["errors"] / r64["totalBits"], r32["errors"] / r32["totalBits"]
    assert abs(ber32 - ber64) <= 0.15 * ber64
    assert r32["uncodedErrors"] / r32["totalBits"] == pytest.approx(r64["uncodedErrors"] / r64["totalBits"], rel=0.01)
    # same frame range: the fp32 front end sees the same channel, so raw errors agree almost exactly
    same = capi.Decoder(code_h, cases.cfg_for(variant, code=code, num_iterations=10, precision=abi.PREC_F32)).simulate(snr, R, 2026, 0, 20000).counters
    ref = capi.Decoder(code_h, cases.cfg_for(variant, code=code, num_iterations=10)).simulate(snr, R, 2026, 0, 20000).counters
    assert abs(same["uncodedErrors"] - ref["uncodedErrors"]) <= 2
    assert abs(same["wordErrors"] - ref["wordErrors"]) <= max(3, 0.05 * ref["wordErrors"])


def test_f16x2_fer_inside_fp64_confidence_interval():
    F = 400000
    code_h = capi.Code(code_path("802_3_H"))
    R = 0.8413
    r64 = capi.Decoder(code_h, cases.cfg_for("decodeNormalizedMinSum", code="802_3_H", num_iterations=10)).simulate(3.6, R, 2026, 0, F).counters
    rh = capi.Decoder(code_h, cases.cfg_for("decodeNormalizedMinSum", code="802_3_H", num_iterations=10,
                                              precision=abi.PREC_F16X2)).simulate(3.6, R, 2026, F, F + 1).counters
    lo, hi = _wilson(r64["wordErrors"], F)
    mid, half = (lo + hi) / 2, (hi - lo) / 2 * math.sqrt(2)
    assert rh["totalWords"] == F + 1
    assert mid - half <= rh["wordErrors"] / (F + 1) <= mid + half, (r64, rh)
    assert abs(rh["errors"] / rh["totalBits"] - r64["errors"] / r64["totalBits"]) <= 0.15 * r64["errors"] / r64["totalBits"]
