"""ctypes binding of the C ABI (include/ldpc_gpu.h -> _build/libldpc_gpu.so).

The library is the product; this file only marshals numpy arrays into its plain-pointer
structs.  Loading fails loudly when the library has not been built, and every compute entry
point fails with ERR_CUDA when no GPU is visible: there is no CPU path to fall back to.
"""
import ctypes as C
import os

import numpy as np

from . import abi

HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("LDPC_GPU_LIB", os.path.join(HERE, "_build", "libldpc_gpu.so"))

# every symbol include/ldpc_gpu.h declares
EXPORTS = [
    "ldpc_gpu_version", "ldpc_gpu_last_error", "ldpc_gpu_init", "ldpc_gpu_shutdown", "ldpc_gpu_device_count",
    "ldpc_gpu_code_create", "ldpc_gpu_code_load_alist", "ldpc_gpu_code_load_alist_transposed", "ldpc_gpu_code_dims", "ldpc_gpu_code_destroy",
    "ldpc_gpu_code_random_codewords",
    "ldpc_gpu_decoder_cfg_default", "ldpc_gpu_iter_hist_len", "ldpc_gpu_decoder_create", "ldpc_gpu_decoder_destroy",
    "ldpc_gpu_decoder_set_codewords", "ldpc_gpu_decode_batch", "ldpc_gpu_simulate", "ldpc_gpu_redecode_stats", "ldpc_gpu_replay_frame", "ldpc_gpu_channel_dump",
    "ldpc_gpu_philox4x32", "ldpc_gpu_last_timing", "ldpc_gpu_decoder_stats", "ldpc_gpu_decoder_geometry",
    "ldpc_gpu_comm_unique_id", "ldpc_gpu_comm_init", "ldpc_gpu_comm_destroy", "ldpc_gpu_allreduce_counters",
    "ldpc_gpu_nb_code_create", "ldpc_gpu_nb_code_load_alist", "ldpc_gpu_nb_code_dims", "ldpc_gpu_nb_code_destroy",
    "ldpc_gpu_nb_decoder_create", "ldpc_gpu_nb_decoder_destroy", "ldpc_gpu_nb_decode_batch", "ldpc_gpu_nb_simulate",
]


class LdpcGpuError(RuntimeError):
    def __init__(self, code, msg):
        super().__init__("ldpc_gpu error %d: %s" % (code, msg))
        self.code = code


_lib = None


def lib():
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise ImportError("%s is missing: run `python -c 'import __graft_entry__ as g; g.build()'` "
                              "(the CUDA library is the only implementation; there is no fallback)" % LIB_PATH)
        L = C.CDLL(LIB_PATH)
        L.ldpc_gpu_last_error.restype = C.c_char_p
        L.ldpc_gpu_code_create.argtypes = [C.c_int] * 4 + [C.c_void_p] * 4 + [C.POINTER(C.c_void_p)]
        L.ldpc_gpu_code_load_alist.argtypes = [C.c_char_p, C.POINTER(C.c_void_p)]
        L.ldpc_gpu_code_load_alist_transposed.argtypes = [C.c_char_p, C.POINTER(C.c_void_p)]
        L.ldpc_gpu_code_dims.argtypes = [C.c_void_p] + [C.POINTER(C.c_int)] * 5
        L.ldpc_gpu_code_destroy.argtypes = [C.c_void_p]
        L.ldpc_gpu_code_random_codewords.argtypes = [C.c_void_p, C.c_uint64, C.c_int64, C.c_void_p, C.POINTER(C.c_int)]
        L.ldpc_gpu_decoder_cfg_default.argtypes = [C.c_int, C.POINTER(abi.DecoderCfg)]
        L.ldpc_gpu_decoder_create.argtypes = [C.c_void_p, C.POINTER(abi.DecoderCfg), C.c_int, C.POINTER(C.c_void_p)]
        L.ldpc_gpu_decoder_destroy.argtypes = [C.c_void_p]
        L.ldpc_gpu_decoder_set_codewords.argtypes = [C.c_void_p, C.c_void_p, C.c_int64]
        L.ldpc_gpu_decode_batch.argtypes = [C.c_void_p, C.POINTER(abi.Channel), C.POINTER(abi.Batch), C.POINTER(abi.Counters)]
        L.ldpc_gpu_simulate.argtypes = [C.c_void_p, C.POINTER(abi.Channel), C.POINTER(abi.SimArgs), C.POINTER(abi.Counters)]
        L.ldpc_gpu_redecode_stats.argtypes = [C.c_void_p, C.POINTER(abi.Channel), C.POINTER(abi.SimArgs), C.c_int32, C.c_void_p,
                                              C.POINTER(abi.Counters)]
        L.ldpc_gpu_replay_frame.argtypes = [C.c_void_p, C.POINTER(abi.Channel), C.c_uint64, C.c_int64, C.c_int32, C.c_void_p, C.c_void_p,
                                            C.POINTER(C.c_int32), C.POINTER(C.c_int32)]
        L.ldpc_gpu_channel_dump.argtypes = [C.c_void_p, C.POINTER(abi.Channel), C.c_uint64, C.c_int64, C.c_int64,
                                            C.c_void_p, C.c_void_p, C.c_int64]
        L.ldpc_gpu_philox4x32.argtypes = [C.c_void_p] * 3
        L.ldpc_gpu_last_timing.argtypes = [C.c_void_p, C.POINTER(C.c_double), C.POINTER(C.c_int64)]
        L.ldpc_gpu_decoder_stats.argtypes = [C.c_void_p, C.POINTER(C.c_int64), C.POINTER(C.c_int32)]
        L.ldpc_gpu_decoder_geometry.argtypes = [C.c_void_p] + [C.POINTER(C.c_int)] * 4
        L.ldpc_gpu_init.argtypes = [C.c_void_p, C.c_int]
        L.ldpc_gpu_comm_unique_id.argtypes = [C.c_void_p]
        L.ldpc_gpu_comm_init.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_int]
        L.ldpc_gpu_allreduce_counters.argtypes = [C.POINTER(abi.Counters), C.c_int, C.POINTER(abi.DecoderCfg)]
        L.ldpc_gpu_iter_hist_len.argtypes = [C.POINTER(abi.DecoderCfg)]
        L.ldpc_gpu_nb_code_load_alist.argtypes = [C.c_char_p, C.POINTER(C.c_void_p)]
        L.ldpc_gpu_nb_code_dims.argtypes = [C.c_void_p] + [C.POINTER(C.c_int)] * 4
        L.ldpc_gpu_nb_code_destroy.argtypes = [C.c_void_p]
        L.ldpc_gpu_nb_decoder_create.argtypes = [C.c_void_p, C.c_int, C.c_int, C.POINTER(C.c_void_p)]
        L.ldpc_gpu_nb_decoder_destroy.argtypes = [C.c_void_p]
        L.ldpc_gpu_nb_decode_batch.argtypes = [C.c_void_p, C.POINTER(abi.Channel), C.c_int64, C.c_void_p, C.c_void_p, C.c_void_p, C.POINTER(abi.Counters)]
        L.ldpc_gpu_nb_simulate.argtypes = [C.c_void_p, C.POINTER(abi.Channel), C.c_uint64, C.c_int64, C.c_int64, C.POINTER(abi.Counters), C.POINTER(C.c_double)]
        _lib = L
    return _lib


def check(rc):
    if rc != 0:
        raise LdpcGpuError(rc, lib().ldpc_gpu_last_error().decode(errors="replace"))


def _ptr(a):
    return None if a is None else a.ctypes.data_as(C.c_void_p)


class Result(dict):
    __getattr__ = dict.__getitem__


class Code:
    """Parity-check matrix handle (reference: alist_struct, inc/alist.h:21-36)."""

    def __init__(self, alist_path=None, arrays=None, transposed=False):
        self.h = C.c_void_p()
        if alist_path is not None and transposed:
            check(lib().ldpc_gpu_code_load_alist_transposed(os.fsencode(alist_path), C.byref(self.h)))
        elif alist_path is not None:
            check(lib().ldpc_gpu_code_load_alist(os.fsencode(alist_path), C.byref(self.h)))
        else:
            N, M, dv, dc, num_n, num_m, nl, ml = arrays
            num_n = np.ascontiguousarray(num_n, np.int32); num_m = np.ascontiguousarray(num_m, np.int32)
            nl = np.ascontiguousarray(nl, np.int32); ml = np.ascontiguousarray(ml, np.int32)
            check(lib().ldpc_gpu_code_create(N, M, dv, dc, _ptr(num_n), _ptr(num_m), _ptr(nl), _ptr(ml), C.byref(self.h)))
        v = [C.c_int() for _ in range(5)]
        check(lib().ldpc_gpu_code_dims(self.h, *[C.byref(x) for x in v]))
        self.N, self.M, self.E, self.dv_max, self.dc_max = [x.value for x in v]

    def random_codewords(self, seed, n):
        """[n][N] bytes 0/1 with H c = 0 (host-side GF(2) encoder); also sets self.rank."""
        out = np.zeros((n, self.N), np.uint8)
        rk = C.c_int()
        check(lib().ldpc_gpu_code_random_codewords(self.h, seed, n, _ptr(out), C.byref(rk)))
        self.rank = rk.value
        return out

    def __del__(self):
        try:
            if self.h:
                lib().ldpc_gpu_code_destroy(self.h)
                self.h = None
        except Exception:
            pass


class Decoder:
    """One decoder variant bound to one GPU (reference: one Makefile binary + its parameter globals)."""

    def __init__(self, code, cfg, device=0):
        self.code, self.cfg, self.N = code, cfg, code.N
        self.h = C.c_void_p()
        check(lib().ldpc_gpu_decoder_create(code.h, C.byref(cfg), device, C.byref(self.h)))

    def __del__(self):
        try:
            if self.h:
                lib().ldpc_gpu_decoder_destroy(self.h)
                self.h = None
        except Exception:
            pass

    def geometry(self):
        v = [C.c_int() for _ in range(4)]
        check(lib().ldpc_gpu_decoder_geometry(self.h, *[C.byref(x) for x in v]))
        return dict(zip(("grid", "block", "smem_bytes", "ctas_per_sm"), [x.value for x in v]))

    def stats(self):
        """(exact_lattice, redo_frames) of the LDPC_GPU_PREC_F16X2 lattice kernel (csrc/ldpc_ms_x2.cuh)."""
        r, x = C.c_int64(), C.c_int32()
        check(lib().ldpc_gpu_decoder_stats(self.h, C.byref(r), C.byref(x)))
        return bool(x.value), r.value

    def last_timing(self):
        ms, n = C.c_double(), C.c_int64()
        check(lib().ldpc_gpu_last_timing(self.h, C.byref(ms), C.byref(n)))
        return ms.value, n.value

    def set_codewords(self, bits01):
        if bits01 is None:
            check(lib().ldpc_gpu_decoder_set_codewords(self.h, None, 0))
        else:
            a = np.ascontiguousarray(bits01, np.uint8)
            assert a.ndim == 2 and a.shape[1] == self.N
            check(lib().ldpc_gpu_decoder_set_codewords(self.h, _ptr(a), a.shape[0]))

    def _counters(self):
        cnt = abi.Counters()
        h = Result(error_weight_hist=np.zeros(self.N, np.int64), iter_hist=np.zeros(abi.iter_hist_len(self.cfg), np.int64),
                   phase_hist=np.zeros(max(1, self.cfg.maxphase), np.int64))
        cnt.error_weight_hist = h.error_weight_hist.ctypes.data_as(C.POINTER(C.c_int64))
        cnt.iter_hist = h.iter_hist.ctypes.data_as(C.POINTER(C.c_int64))
        cnt.phase_hist = h.phase_hist.ctypes.data_as(C.POINTER(C.c_int64))
        return cnt, h

    def decode(self, snr_db, R, y, noise=None, noise_rows=0, codeword=None, qpointer0=None, count=True,
               want_soft=True, y_dtype=abi.DT_F64, outputs=True):
        """Host-memory parity entry.  y: [F][N] raw channel samples."""
        N = self.N
        y = np.ascontiguousarray(y, dtype={abi.DT_F64: np.float64, abi.DT_F32: np.float32, abi.DT_F16: np.float16, abi.DT_Q8: np.int8, abi.DT_QP: np.uint8}[y_dtype])
        F = y.shape[0]
        out = Result(bits=np.zeros((F, (N + 7) // 8), np.uint8) if outputs else None,
                     iters=np.zeros(F, np.int32) if outputs else None,
                     soft=np.zeros((F, N), np.float64 if y_dtype == abi.DT_F64 else np.float32) if (want_soft and outputs) else None,
                     errors=np.zeros(F, np.int32) if outputs else None,
                     flags=np.zeros(F, np.uint8) if outputs else None)
        b = abi.Batch()
        b.n_frames, b.mem, b.y_dtype = F, abi.MEM_HOST, y_dtype
        b.y = _ptr(y)
        noise = None if noise is None else np.ascontiguousarray(noise, np.float64)
        b.noise, b.noise_rows = _ptr(noise), int(noise_rows)
        codeword = None if codeword is None else np.ascontiguousarray(codeword, np.uint8)
        b.codeword = _ptr(codeword)
        qpointer0 = None if qpointer0 is None else np.ascontiguousarray(qpointer0, np.int32)
        b.qpointer0 = _ptr(qpointer0)
        b.out_bits, b.out_iters, b.out_soft = _ptr(out.bits), _ptr(out.iters), _ptr(out.soft)
        b.out_errors, b.out_flags = _ptr(out.errors), _ptr(out.flags)
        ch = abi.Channel(snr_db, R)
        cnt, hist = self._counters()
        check(lib().ldpc_gpu_decode_batch(self.h, C.byref(ch), C.byref(b), C.byref(cnt) if count else None))
        out["counters"] = cnt.as_dict()
        out.update(hist)
        if outputs:
            out["d"] = np.unpackbits(out.bits, axis=1, bitorder="little")[:, :N]
        return out

    def decode_raw(self, snr_db, R, batch, counters=None):
        """Pass a prepared abi.Batch straight through (device pointers, pinned buffers, ...)."""
        ch = abi.Channel(snr_db, R)
        check(lib().ldpc_gpu_decode_batch(self.h, C.byref(ch), C.byref(batch), None if counters is None else C.byref(counters)))

    def simulate(self, snr_db, R, seed, frame_begin, n_frames, stop_errors=0, stop_word_errors=0, poll_frames=0):
        ch = abi.Channel(snr_db, R)
        a = abi.SimArgs(seed, frame_begin, n_frames, stop_errors, stop_word_errors, poll_frames)
        cnt, hist = self._counters()
        check(lib().ldpc_gpu_simulate(self.h, C.byref(ch), C.byref(a), C.byref(cnt)))
        r = Result(counters=cnt.as_dict())
        r.update(hist)
        r["_cnt"] = cnt
        return r

    def redecode_stats(self, snr_db, R, seed, frame_begin, n_frames, n_redecodes):
        """[n_frames][n_redecodes] error weights of n_redecodes decodes of every frame (same channel, fresh decoder noise)
        and the counters accumulated over all of them."""
        ch = abi.Channel(snr_db, R)
        a = abi.SimArgs(seed, frame_begin, n_frames, 0, 0, 0)
        cnt, hist = self._counters()
        out = np.zeros((n_frames, n_redecodes), np.int32)
        check(lib().ldpc_gpu_redecode_stats(self.h, C.byref(ch), C.byref(a), int(n_redecodes), _ptr(out), C.byref(cnt)))
        return out, cnt.as_dict()

    def replay_frame(self, snr_db, R, seed, frame_id, max_rows=None):
        """Per-iteration trace of one seed-addressed frame: (d [rows][N] 0/1, syn [rows][M] 0/1, rows executed, final error weight)."""
        T = self.cfg.num_iterations
        max_rows = T if max_rows is None else max_rows
        M = self.code.M
        td = np.zeros((max_rows, (self.N + 7) // 8), np.uint8)
        ts = np.zeros((max_rows, (M + 7) // 8), np.uint8)
        n, e = C.c_int32(), C.c_int32()
        ch = abi.Channel(snr_db, R)
        check(lib().ldpc_gpu_replay_frame(self.h, C.byref(ch), seed, frame_id, max_rows, _ptr(td), _ptr(ts), C.byref(n), C.byref(e)))
        k = min(n.value, max_rows)
        return (np.unpackbits(td[:k], axis=1, bitorder="little")[:, :self.N], np.unpackbits(ts[:k], axis=1, bitorder="little")[:, :M], n.value, e.value)

    def channel_dump(self, snr_db, R, seed, frame_begin, n_frames, noise_rows=None):
        ch = abi.Channel(snr_db, R)
        y = np.zeros((n_frames, self.N), np.float64)
        if noise_rows is None:
            noise_rows = abi.noise_rows_needed(self.cfg)
        noise = None
        if self.cfg.kind == abi.KIND_NGDBF_HW:
            noise = np.zeros((n_frames, abi.HW_QBUF), np.float64)
        elif self.cfg.kind == abi.KIND_NGDBF_SC:
            noise_rows = abi.sc_noise_len(self.cfg, self.N)
            noise = np.zeros((n_frames, noise_rows), np.float64)
        elif noise_rows:
            noise = np.zeros((n_frames, noise_rows, self.N), np.float64)
        check(lib().ldpc_gpu_channel_dump(self.h, C.byref(ch), seed, frame_begin, n_frames, _ptr(y), _ptr(noise), noise_rows))
        return y, noise


class NbCode:
    """Non-binary GF(q) parity-check matrix (format of SystemC/NB-LDPC/src/alist.cpp:23-56)."""

    def __init__(self, alist_path):
        self.h = C.c_void_p()
        check(lib().ldpc_gpu_nb_code_load_alist(os.fsencode(alist_path), C.byref(self.h)))
        v = [C.c_int() for _ in range(4)]
        check(lib().ldpc_gpu_nb_code_dims(self.h, *[C.byref(x) for x in v]))
        self.N, self.M, self.q, self.E = [x.value for x in v]
        self.m = self.q.bit_length() - 1

    def __del__(self):
        try:
            if self.h:
                lib().ldpc_gpu_nb_code_destroy(self.h)
                self.h = None
        except Exception:
            pass


class NbDecoder:
    """Min-max decoder of a non-binary code on one GPU (csrc/ldpc_nb_kernel.cuh)."""

    def __init__(self, code, num_iterations, device=0):
        self.code, self.T = code, num_iterations
        self.h = C.c_void_p()
        check(lib().ldpc_gpu_nb_decoder_create(code.h, num_iterations, device, C.byref(self.h)))

    def __del__(self):
        try:
            if self.h:
                lib().ldpc_gpu_nb_decoder_destroy(self.h)
                self.h = None
        except Exception:
            pass

    def decode(self, snr_db, R, y):
        """y: [F][N*m] bit samples.  Returns symbols [F][N], iterations [F], counters."""
        y = np.ascontiguousarray(y, np.float64)
        F = y.shape[0]
        sym, it = np.zeros((F, self.code.N), np.uint8), np.zeros(F, np.int32)
        cnt, ch = abi.Counters(), abi.Channel(snr_db, R)
        check(lib().ldpc_gpu_nb_decode_batch(self.h, C.byref(ch), F, _ptr(y), _ptr(sym), _ptr(it), C.byref(cnt)))
        return Result(symbols=sym, iters=it, counters=cnt.as_dict())

    def simulate(self, snr_db, R, seed, frame_begin, n_frames):
        cnt, ch, ms = abi.Counters(), abi.Channel(snr_db, R), C.c_double()
        check(lib().ldpc_gpu_nb_simulate(self.h, C.byref(ch), seed, frame_begin, n_frames, C.byref(cnt), C.byref(ms)))
        return Result(counters=cnt.as_dict(), kernel_ms=ms.value)


def philox4x32(ctr, key):
    c = np.asarray(ctr, np.uint32); k = np.asarray(key, np.uint32); o = np.zeros(4, np.uint32)
    check(lib().ldpc_gpu_philox4x32(_ptr(c), _ptr(k), _ptr(o)))
    return o


def device_count():
    return lib().ldpc_gpu_device_count()
