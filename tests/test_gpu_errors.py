"""Error convention of the C ABI on a live GPU: bad input yields a negative code and a message, never a
crash (the reference has no error returns: a bad alist path dereferences NULL, src/alist.cpp:71-74)."""
import ctypes as C

import numpy as np
import pytest

import cases
from ldpcsimulation_b200 import abi, capi
from oracle.oracle_api import code_path

pytestmark = pytest.mark.gpu


def _batch(F, N, y):
    b = abi.Batch()
    b.n_frames, b.mem, b.y_dtype = F, abi.MEM_HOST, abi.DT_F64
    b.y = y.ctypes.data_as(C.c_void_p)
    return b


def _rc(fn, *a):
    rc = fn(*a)
    return rc, capi.lib().ldpc_gpu_last_error().decode()


def test_decoder_creation_rejects_bad_configurations():
    code = capi.Code(code_path("PEG"))
    for kw, code_want in [(dict(num_iterations=-1), abi.ERR_INVALID_ARG), (dict(precision=7), abi.ERR_INVALID_ARG),
                          (dict(kind=9), abi.ERR_INVALID_ARG)]:
        cfg = cases.cfg_for("decodeMinSum")
        for k, v in kw.items():
            setattr(cfg, k, v)
        with pytest.raises(capi.LdpcGpuError) as e:
            capi.Decoder(code, cfg)
        assert e.value.code == code_want
    with pytest.raises(capi.LdpcGpuError) as e:                    # RNGDBF.cpp has no sample quantiser
        capi.Decoder(code, abi.default_cfg(abi.KIND_GDBF, flags=["redecode", "quantizeSamples"]))
    assert e.value.code == abi.ERR_UNSUPPORTED
    with pytest.raises(capi.LdpcGpuError) as e:                    # NGDBFhw's 2648-entry window needs N < 2648
        capi.Decoder(capi.Code(code_path("4000")), abi.default_cfg(abi.KIND_NGDBF_HW))
    assert e.value.code == abi.ERR_UNSUPPORTED and "2648" in str(e.value)
    with pytest.raises(capi.LdpcGpuError) as e:
        capi.Decoder(code, cases.cfg_for("decodeMinSum"), device=99)
    assert e.value.code == abi.ERR_CUDA


def test_decode_batch_argument_checks():
    L = capi.lib()
    code = capi.Code(code_path("PEG"))
    ch = abi.Channel(3.0, 0.5)
    y = np.ones((2, code.N))
    dec = capi.Decoder(code, cases.cfg_for("decodeMinSum"))
    b = _batch(2, code.N, y); b.y = None
    rc, msg = _rc(L.ldpc_gpu_decode_batch, dec.h, C.byref(ch), C.byref(b), None)
    assert rc == abi.ERR_INVALID_ARG and "batch.y" in msg
    b = _batch(-1, code.N, y)
    assert L.ldpc_gpu_decode_batch(dec.h, C.byref(ch), C.byref(b), None) == abi.ERR_INVALID_ARG
    b = _batch(2, code.N, y); b.y_dtype = 5
    assert L.ldpc_gpu_decode_batch(dec.h, C.byref(ch), C.byref(b), None) == abi.ERR_INVALID_ARG
    b = _batch(2, code.N, y); b.mem = 3
    assert L.ldpc_gpu_decode_batch(dec.h, C.byref(ch), C.byref(b), None) == abi.ERR_INVALID_ARG
    bad = abi.Channel(3.0, 0.0)
    b = _batch(2, code.N, y)
    assert L.ldpc_gpu_decode_batch(dec.h, C.byref(bad), C.byref(b), None) == abi.ERR_INVALID_ARG
    assert L.ldpc_gpu_decode_batch(None, C.byref(ch), C.byref(b), None) == abi.ERR_INVALID_ARG
    # the noisy bit-flipping variants need their raw RNG outputs through the parity entry
    ng = capi.Decoder(code, cases.cfg_for("decodeSMNGDBF"))
    rc, msg = _rc(L.ldpc_gpu_decode_batch, ng.h, C.byref(ch), C.byref(b), None)
    assert rc == abi.ERR_INVALID_ARG and "noise" in msg
    noise = np.zeros((2, 3, code.N))
    b.noise, b.noise_rows = noise.ctypes.data_as(C.c_void_p), 3
    rc, msg = _rc(L.ldpc_gpu_decode_batch, ng.h, C.byref(ch), C.byref(b), None)
    assert rc == abi.ERR_INVALID_ARG and "noise_rows" in msg
    hw = capi.Decoder(capi.Code(code_path("802_3_H")), cases.cfg_for("NGDBFhw"))
    b = _batch(1, 2048, np.ones((1, 2048)))
    assert L.ldpc_gpu_decode_batch(hw.h, C.byref(ch), C.byref(b), None) == abi.ERR_INVALID_ARG


def test_simulate_and_misc_argument_checks():
    L = capi.lib()
    code = capi.Code(code_path("PEG"))
    cfg = cases.cfg_for("decodeMinSum")
    dec = capi.Decoder(code, cfg)
    ch = abi.Channel(3.0, 0.5)
    cnt = abi.Counters()
    a = abi.SimArgs(1, 0, -5, 0, 0, 0)
    assert L.ldpc_gpu_simulate(dec.h, C.byref(ch), C.byref(a), C.byref(cnt)) == abi.ERR_INVALID_ARG
    a = abi.SimArgs(1, 0, 10, 0, 0, 0)
    assert L.ldpc_gpu_simulate(dec.h, C.byref(ch), C.byref(a), None) == abi.ERR_INVALID_ARG
    assert L.ldpc_gpu_simulate(dec.h, C.byref(ch), C.byref(a), C.byref(cnt)) == 0 and cnt.totalWords == 10
    assert L.ldpc_gpu_simulate(dec.h, C.byref(ch), C.byref(a), C.byref(cnt)) == 0 and cnt.totalWords == 20   # counters are ADDED
    assert L.ldpc_gpu_allreduce_counters(C.byref(cnt), code.N, C.byref(cfg)) == abi.ERR_COMM                        # no communicator
    assert L.ldpc_gpu_init((C.c_int * 1)(42), 1) == abi.ERR_INVALID_ARG
    assert L.ldpc_gpu_init(None, 0) == 0
    assert L.ldpc_gpu_decoder_destroy(None) == 0 and L.ldpc_gpu_code_destroy(None) == 0
    g = dec.geometry()
    assert g["grid"] % 148 == 0 and g["block"] % 32 == 0


def test_redecode_stats_argument_checks():
    L = capi.lib()
    code = capi.Code(code_path("PEG"))
    ch = abi.Channel(4.0, 0.5)
    out = np.zeros((4, 3), np.int32)
    a = abi.SimArgs(1, 0, 4, 0, 0, 0)
    ms = capi.Decoder(code, cases.cfg_for("decodeMinSum"))
    rc, msg = _rc(L.ldpc_gpu_redecode_stats, ms.h, C.byref(ch), C.byref(a), 3, out.ctypes.data_as(C.c_void_p), None)
    assert rc == abi.ERR_UNSUPPORTED and "noise" in msg                     # a decoder without its own randomness has nothing to re-decode
    ng = capi.Decoder(code, cases.cfg_for("decodeSMNGDBF", num_iterations=10))
    assert L.ldpc_gpu_redecode_stats(ng.h, C.byref(ch), C.byref(a), 0, out.ctypes.data_as(C.c_void_p), None) == abi.ERR_INVALID_ARG
    assert L.ldpc_gpu_redecode_stats(ng.h, C.byref(ch), C.byref(a), 3, None, None) == abi.ERR_INVALID_ARG
    assert L.ldpc_gpu_redecode_stats(ng.h, C.byref(ch), C.byref(a), 3, out.ctypes.data_as(C.c_void_p), None) == 0     # counters are optional
    empty = abi.SimArgs(1, 0, 0, 0, 0, 0)
    assert L.ldpc_gpu_redecode_stats(ng.h, C.byref(ch), C.byref(empty), 3, out.ctypes.data_as(C.c_void_p), None) == 0
