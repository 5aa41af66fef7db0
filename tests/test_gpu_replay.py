"""ldpc_gpu_replay_frame (replay half): the per-iteration trace of one seed-addressed frame equals the
oracle's state after t = 1, 2, ... iterations on the dumped samples and noise (the format of src/replayGDBF.cpp:312-314,368-373:
decisions after the step, syndromes the step started from)."""
import numpy as np
import pytest

import cases
from ldpcsimulation_b200 import abi, capi
from oracle.oracle_api import Oracle, code_path

pytestmark = pytest.mark.gpu


def _H(path):
    t = list(map(int, open(path).read().split()))
    N, M, dv, dc = t[0], t[1], t[2], t[3]
    ml = np.array(t[4 + N + M + N * dv:4 + N + M + N * dv + M * dc]).reshape(M, dc)
    H = np.zeros((M, N), np.uint8)
    for j in range(M):
        for v in ml[j]:
            if v > 0:
                H[j, v - 1] = 1
    return H


@pytest.mark.parametrize("variant,code_name,T", [("decodeSMNGDBF", "PEG", 40), ("decodeGDBF", "PEG", 25), ("decodeRSMNGDBF", "802_3_H", 30),
                                                 ("NGDBFhw", "802_3_H", 60), ("decodeMinSum", "PEG", 8), ("decodeOffsetMinSum", "802_3_H", 6)])
def test_replay_trace_equals_oracle_states(variant, code_name, T):
    R, snr = cases.operating_point(variant, code_name)
    snr -= 0.6                                            # frames that take a while (and some that never finish)
    cfg = cases.cfg_for(variant, code=code_name, num_iterations=T)
    code = capi.Code(code_path(code_name))
    dec = capi.Decoder(code, cfg)
    orc = Oracle(code_name)
    H = _H(code_path(code_name))
    for frame in (3, 1000003):
        d, syn, rows, err = dec.replay_frame(snr, R, 77, frame)
        y, noise = dec.channel_dump(snr, R, 77, frame, 1)
        nrows = abi.noise_rows_needed(cfg)
        # the oracle's decisions after t iterations, t = 0 .. rows (smoothing / further phases are not part of the trace)
        cfg_t = cases.cfg_for(variant, code=code_name, num_iterations=T)
        cfg_t.flags &= ~abi.F_OUTPUTSMOOTHING
        if cfg_t.flags & abi.F_REDECODE:
            cfg_t.maxphase = 1
        states = []
        for t in range(0, rows + 1):
            cfg_t.num_iterations = t
            if t == 0 and cfg.kind in (abi.KIND_GDBF, abi.KIND_NGDBF_HW):
                states.append((y[0] <= 0).astype(np.uint8))              # r: the channel's hard decisions (decodeGDBF.cpp:259-267)
                continue
            nz = None if noise is None else (noise if cfg.kind == abi.KIND_NGDBF_HW else noise[:, :max(1, t * abi.gdbf_rows_per_step(cfg.flags))])
            o = orc.decode(cfg_t, snr, R, y, nz, 0 if nz is None or cfg.kind == abi.KIND_NGDBF_HW else nz.shape[1])
            states.append(o.d[0])
        full = orc.decode(cfg_t, snr, R, y, noise, nrows if cfg.kind == abi.KIND_GDBF else 0)      # T = rows here
        assert rows == (full.iters[0] if cfg.kind in (abi.KIND_GDBF, abi.KIND_NGDBF_HW) else T)
        for t in range(rows):
            assert np.array_equal(d[t], states[t + 1]), (variant, frame, t)
            assert np.array_equal(syn[t], (H @ states[t]) & 1), (variant, frame, t)
        assert err == int(states[rows].sum())                              # all-zero codeword
