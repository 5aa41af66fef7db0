"""The C-ABI library loads, exports every symbol include/ldpc_gpu.h declares, and refuses to compute
without a GPU (no CPU fallback).  No compute calls here."""
import ctypes as C
import os
import re

import pytest

from ldpcsimulation_b200 import abi, capi
from oracle.oracle_api import code_path

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def header_symbols():
    text = open(os.path.join(ROOT, "include", "ldpc_gpu.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(ldpc_gpu_[a-z0-9_]+)\s*\(", text)))


def test_library_exports_every_declared_symbol():
    L = capi.lib()
    syms = header_symbols()
    assert sorted(capi.EXPORTS) == syms
    for s in syms:
        assert hasattr(L, s), s
    assert L.ldpc_gpu_version() >= 100


def test_struct_layouts_match_header():
    # sizes the C compiler gives the header's structs
    import subprocess, tempfile, textwrap
    src = textwrap.dedent('''
        #include <stdio.h>
        #include <stddef.h>
        #include "ldpc_gpu.h"
        int main(void) { printf("%zu %zu %zu %zu %zu %zu %zu\\n", sizeof(ldpc_gpu_decoder_cfg), sizeof(ldpc_gpu_channel),
            sizeof(ldpc_gpu_counters), sizeof(ldpc_gpu_batch), sizeof(ldpc_gpu_sim_args),
            offsetof(ldpc_gpu_decoder_cfg, MAXLLR), offsetof(ldpc_gpu_batch, out_flags)); return 0; }''')
    with tempfile.TemporaryDirectory() as t:
        open(os.path.join(t, "s.c"), "w").write(src)
        subprocess.check_call(["gcc", "-I" + os.path.join(ROOT, "include"), os.path.join(t, "s.c"), "-o", os.path.join(t, "s")])
        got = [int(x) for x in subprocess.check_output([os.path.join(t, "s")]).split()]
    want = [C.sizeof(abi.DecoderCfg), C.sizeof(abi.Channel), C.sizeof(abi.Counters), C.sizeof(abi.Batch), C.sizeof(abi.SimArgs),
            abi.DecoderCfg.MAXLLR.offset, abi.Batch.out_flags.offset]
    assert got == want


def test_defaults_match_reference_globals():
    # src/decodeGDBF.cpp:48-56, src/NGDBFhw.cpp:48-57, src/decodeBP.cpp:58
    c = abi.DecoderCfg()
    capi.check(capi.lib().ldpc_gpu_decoder_cfg_default(abi.KIND_GDBF, C.byref(c)))
    assert (c.lambda_, c.alpha, c.Ymax, c.windowsize, c.noiseScale, c.NQ, c.Tswitch, c.maxphase) == (0.991, 2.25, 2.25, 64, 1.0, 16, 0, 7)
    capi.check(capi.lib().ldpc_gpu_decoder_cfg_default(abi.KIND_NGDBF_HW, C.byref(c)))
    assert (c.num_iterations, c.w, c.Ymax, c.noiseScale, c.maxphase, c.NQ, c.theta0) == (600, 0.185, 1.625, 0.95, 1, 5, -0.525)
    capi.check(capi.lib().ldpc_gpu_decoder_cfg_default(abi.KIND_BP, C.byref(c)))
    assert c.MAXLLR == 20
    p = abi.default_cfg(abi.KIND_NGDBF_HW)
    assert (p.num_iterations, p.w, p.Ymax, p.noiseScale, p.maxphase, p.NQ, p.theta0) == (600, 0.185, 1.625, 0.95, 1, 5, -0.525)


def test_no_cpu_fallback():
    """Without a GPU every compute entry point must fail loudly with ERR_CUDA."""
    if capi.device_count() > 0:
        pytest.skip("a GPU is visible")
    code = capi.Code(code_path("PEG"))
    with pytest.raises(capi.LdpcGpuError) as e:
        capi.Decoder(code, abi.default_cfg(abi.KIND_MINSUM))
    assert e.value.code == abi.ERR_CUDA and "no CPU fallback" in str(e.value)
    with pytest.raises(capi.LdpcGpuError):
        capi.philox4x32((0, 0, 0, 0), (0, 0))
    assert capi.lib().ldpc_gpu_init(None, 0) == abi.ERR_CUDA


def test_product_never_imports_the_oracle():
    pkg = os.path.join(ROOT, "ldpcsimulation_b200")
    for dp, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".cpp", ".h")):
                text = open(os.path.join(dp, f)).read()
                assert "oracle_api" not in text and "ldpc_oracle" not in text and "oracle/" not in text, os.path.join(dp, f)
