"""Committed outputs of the reference's own object code (tests/golden, made by tools/make_golden.py)
reproduced by the oracle (CPU) and by the CUDA path (GPU), bit for bit."""
import glob
import os

import numpy as np
import pytest

import cases
from ldpcsimulation_b200 import abi
from oracle.oracle_api import Oracle, code_path

GOLDEN = sorted(glob.glob(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "*.npz")))


def _load(path):
    import sys
    sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tools"))
    import make_golden
    variant, code = os.path.basename(path)[:-4].split("__")
    g = np.load(path)
    cfg = cases.cfg_for(variant, code=code)
    orc = Oracle(code)
    R, snr, y, noise, cw = make_golden.golden_inputs(orc, cfg, variant, code)
    assert (R, snr) == (float(g["R"]), float(g["snr"]))
    assert make_golden.sha(y) == str(g["y_sha256"]) and make_golden.sha(noise) == str(g["noise_sha256"]), \
        "the counter-based channel generator drifted: golden inputs can no longer be regenerated"
    qp = g["qpointer0"] if cfg.kind == abi.KIND_NGDBF_HW else None
    return variant, code, g, cfg, orc, R, snr, y, noise, cw, qp


def _check(out, g, cfg, soft_tol=None):
    assert np.array_equal(out.bits, g["bits"])
    assert np.array_equal(out.iters, g["iters"])
    assert np.array_equal(out.errors, g["errors"])
    assert np.array_equal(out.flags, g["flags"])
    assert [out.counters[k] for k in abi.Counters.SCALARS] == g["counters"].tolist()
    assert np.array_equal(out.error_weight_hist, g["error_weight_hist"])
    assert np.array_equal(out.iter_hist, g["iter_hist"])
    assert np.array_equal(out.phase_hist, g["phase_hist"])
    if g["soft"].size:
        if soft_tol:
            np.testing.assert_allclose(out.soft, g["soft"], rtol=soft_tol, atol=1e-12)
        else:
            assert np.array_equal(out.soft, g["soft"])


def test_golden_set_is_complete():
    names = {os.path.basename(p)[:-4] for p in GOLDEN}
    for v in cases.VARIANTS:
        assert ("%s__802_3_H" % v in names) or v == "decodeBP"
        assert ("%s__PEG" % v in names) or v == "NGDBFhw"


@pytest.mark.parametrize("path", GOLDEN, ids=[os.path.basename(p)[:-4] for p in GOLDEN])
def test_oracle_reproduces_reference_outputs(path):
    variant, code, g, cfg, orc, R, snr, y, noise, cw, qp = _load(path)
    _check(orc.decode(cfg, snr, R, y, noise, abi.noise_rows_needed(cfg), cw, qpointer0=qp), g, cfg)


@pytest.mark.gpu
@pytest.mark.parametrize("path", GOLDEN, ids=[os.path.basename(p)[:-4] for p in GOLDEN])
def test_cuda_reproduces_reference_outputs(path):
    from ldpcsimulation_b200 import capi
    variant, code, g, cfg, orc, R, snr, y, noise, cw, qp = _load(path)
    dec = capi.Decoder(capi.Code(code_path(code)), cfg)
    out = dec.decode(snr, R, y, noise, abi.noise_rows_needed(cfg), cw, qpointer0=qp)
    _check(out, g, cfg, soft_tol=1e-9 if cfg.kind == abi.KIND_BP else None)
