"""ctypes access to the oracle (C restatement) and to oracle/_ref (the reference's own object code).

TEST INFRASTRUCTURE ONLY: imported by tests/, __graft_entry__.smoke() and bench.py's CPU-baseline
legs.  Nothing under ldpcsimulation_b200/ imports this.
"""
import ctypes as C
import os

import numpy as np

from ldpcsimulation_b200 import abi

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
CODES = os.path.join(ROOT, "codes")


def code_path(name):
    """'802_3_H' -> codes/802_3/802_3_H.alist etc."""
    table = {
        "PEG": "PEGReg504x1008/PEGReg504x1008.alist",
        "802_3_H": "802_3/802_3_H.alist",
        "802_3": "802_3/802_3.alist",
        "4000": "4000.2000.4.244/4000.2000.4.244.alist",
        "4376": "4376.282.4.9598/4376.282.4.9598.alist",
        "dvbs2": "dvbs2_1_2/dvbs2_1_2.alist",
    }
    return os.path.join(CODES, table.get(name, name))


def load_codewords(path, limit=None):
    """data.enc: one codeword per line, N chars '0'/'1' (src/decodeMinSum.cpp:195-211)."""
    rows = []
    with open(path) as f:
        for line in f:
            line = line.strip()
            if not line:
                continue
            rows.append(np.frombuffer(line.encode(), dtype=np.uint8) - ord("0"))
            if limit and len(rows) >= limit:
                break
    return np.ascontiguousarray(np.stack(rows).astype(np.uint8))


def _ptr(a):
    return None if a is None else a.ctypes.data_as(C.c_void_p)


class Result(dict):
    __getattr__ = dict.__getitem__


def make_batch(N, y, noise=None, noise_rows=0, codeword=None, qpointer0=None, want_soft=True, y_dtype=abi.DT_F64,
               keep=None):
    """Allocate output arrays and fill an abi.Batch for F frames of HOST data."""
    y = np.ascontiguousarray(y, dtype=np.float64 if y_dtype == abi.DT_F64 else np.float32)
    F = y.shape[0]
    out = Result(
        bits=np.zeros((F, (N + 7) // 8), np.uint8), iters=np.zeros(F, np.int32),
        soft=np.zeros((F, N), y.dtype) if want_soft else None,
        errors=np.zeros(F, np.int32), flags=np.zeros(F, np.uint8))
    b = abi.Batch()
    b.n_frames, b.mem, b.y_dtype = F, abi.MEM_HOST, y_dtype
    b.y = _ptr(y)
    if noise is not None:
        noise = np.ascontiguousarray(noise, dtype=np.float64)
    b.noise, b.noise_rows = _ptr(noise), int(noise_rows)
    if codeword is not None:
        codeword = np.ascontiguousarray(codeword, dtype=np.uint8)
        assert codeword.shape == (F, N)
    b.codeword = _ptr(codeword)
    if qpointer0 is not None:
        qpointer0 = np.ascontiguousarray(qpointer0, dtype=np.int32)
    b.qpointer0 = _ptr(qpointer0)
    b.out_bits, b.out_iters, b.out_soft = _ptr(out.bits), _ptr(out.iters), _ptr(out.soft)
    b.out_errors, b.out_flags = _ptr(out.errors), _ptr(out.flags)
    if keep is not None:
        keep.extend([y, noise, codeword, qpointer0])
    out["_keep"] = [y, noise, codeword, qpointer0]
    return b, out


def make_counters(N, cfg):
    cnt = abi.Counters()
    h = Result(error_weight_hist=np.zeros(N, np.int64), iter_hist=np.zeros(abi.iter_hist_len(cfg), np.int64),
               phase_hist=np.zeros(max(1, cfg.maxphase), np.int64))
    cnt.error_weight_hist = h.error_weight_hist.ctypes.data_as(C.POINTER(C.c_int64))
    cnt.iter_hist = h.iter_hist.ctypes.data_as(C.POINTER(C.c_int64))
    cnt.phase_hist = h.phase_hist.ctypes.data_as(C.POINTER(C.c_int64))
    return cnt, h


def unpack_bits(bits, N):
    return np.unpackbits(bits, axis=1, bitorder="little")[:, :N]


# ------------------------------------------------------------------------------------------------
class Oracle:
    """The C restatement (oracle/ldpc_oracle.c)."""

    _lib = None

    @classmethod
    def lib(cls):
        if cls._lib is None:
            path = os.path.join(HERE, "_build", "libldpc_oracle.so")
            if not os.path.exists(path):
                import importlib.util
                spec = importlib.util.spec_from_file_location("build_ref", os.path.join(HERE, "build_ref.py"))
                m = importlib.util.module_from_spec(spec)
                spec.loader.exec_module(m)
                m.build_restatement()
            L = C.CDLL(path)
            L.oracle_code_load_alist.restype = C.c_void_p
            L.oracle_code_load_alist.argtypes = [C.c_char_p]
            L.oracle_code_create.restype = C.c_void_p
            L.oracle_code_create.argtypes = [C.c_int] * 4 + [C.c_void_p] * 4
            L.oracle_code_free.argtypes = [C.c_void_p]
            L.oracle_last_error.restype = C.c_char_p
            L.oracle_decode_batch.argtypes = [C.c_void_p, C.POINTER(abi.DecoderCfg), C.POINTER(abi.Channel),
                                              C.POINTER(abi.Batch), C.POINTER(abi.Counters)]
            L.oracle_simulate.argtypes = [C.c_void_p, C.POINTER(abi.DecoderCfg), C.POINTER(abi.Channel),
                                          C.POINTER(abi.SimArgs), C.c_void_p, C.c_int64, C.POINTER(abi.Counters)]
            L.oracle_channel_dump.argtypes = [C.c_void_p, C.POINTER(abi.DecoderCfg), C.POINTER(abi.Channel), C.c_uint64,
                                              C.c_int64, C.c_int64, C.c_void_p, C.c_int64, C.c_void_p, C.c_void_p, C.c_int64]
            L.oracle_philox4x32.argtypes = [C.c_void_p] * 3
            L.oracle_normal4.argtypes = [C.c_uint64, C.c_uint64, C.c_uint32, C.c_uint32, C.c_uint32, C.c_void_p]
            L.oracle_quantize_ms.restype = C.c_double
            L.oracle_quantize_ms.argtypes = [C.c_double] * 3
            L.oracle_quantize_gdbf.restype = C.c_double
            L.oracle_quantize_gdbf.argtypes = [C.c_double, C.c_double, C.c_int]
            L.oracle_hw_pack.argtypes = [C.c_double] * 3
            L.oracle_hw_unpack.argtypes = [C.c_int]
            cls._lib = L
        return cls._lib

    def __init__(self, alist):
        self.path = alist if os.path.exists(alist) else code_path(alist)
        L = self.lib()
        self.h = L.oracle_code_load_alist(self.path.encode())
        if not self.h:
            raise ValueError("oracle: %s: %s" % (self.path, L.oracle_last_error().decode()))

        class _Code(C.Structure):
            _fields_ = [("N", C.c_int), ("M", C.c_int), ("E", C.c_int), ("dv_max", C.c_int), ("dc_max", C.c_int)]
        c = _Code.from_address(self.h)
        self.N, self.M, self.E, self.dv_max, self.dc_max = c.N, c.M, c.E, c.dv_max, c.dc_max

    def __del__(self):
        try:
            if getattr(self, "h", None):
                self.lib().oracle_code_free(self.h)
                self.h = None
        except Exception:
            pass

    def decode(self, cfg, snr_db, R, y, noise=None, noise_rows=0, codeword=None, qpointer0=None, count=True):
        b, out = make_batch(self.N, y, noise, noise_rows, codeword, qpointer0)
        ch = abi.Channel(snr_db, R)
        cnt, hist = make_counters(self.N, cfg)
        rc = self.lib().oracle_decode_batch(self.h, C.byref(cfg), C.byref(ch), C.byref(b), C.byref(cnt) if count else None)
        if rc:
            raise RuntimeError("oracle_decode_batch: %s" % self.lib().oracle_last_error().decode())
        out["counters"] = cnt.as_dict()
        out.update(hist)
        out["d"] = unpack_bits(out.bits, self.N)
        return out

    def channel_dump(self, cfg, snr_db, R, seed, frame_begin, n_frames, codewords=None, noise_rows=None):
        ch = abi.Channel(snr_db, R)
        y = np.zeros((n_frames, self.N), np.float64)
        if noise_rows is None:
            noise_rows = abi.noise_rows_needed(cfg)
        noise = None
        if cfg.kind == abi.KIND_NGDBF_HW:
            noise = np.zeros((n_frames, abi.HW_QBUF), np.float64)
        elif cfg.kind == abi.KIND_NGDBF_SC:
            noise_rows = abi.sc_noise_len(cfg, self.N)
            noise = np.zeros((n_frames, noise_rows), np.float64)
        elif noise_rows:
            noise = np.zeros((n_frames, noise_rows, self.N), np.float64)
        ncw = 0 if codewords is None else len(codewords)
        self.lib().oracle_channel_dump(self.h, C.byref(cfg), C.byref(ch), seed, frame_begin, n_frames,
                                       _ptr(codewords), ncw, _ptr(y), _ptr(noise), noise_rows)
        return y, noise

    def simulate(self, cfg, snr_db, R, seed, frame_begin, n_frames, codewords=None, stop_errors=0, stop_word_errors=0):
        ch = abi.Channel(snr_db, R)
        a = abi.SimArgs(seed, frame_begin, n_frames, stop_errors, stop_word_errors, 0)
        cnt, hist = make_counters(self.N, cfg)
        ncw = 0 if codewords is None else len(codewords)
        rc = self.lib().oracle_simulate(self.h, C.byref(cfg), C.byref(ch), C.byref(a), _ptr(codewords), ncw, C.byref(cnt))
        if rc:
            raise RuntimeError("oracle_simulate: %s" % self.lib().oracle_last_error().decode())
        r = Result(counters=cnt.as_dict())
        r.update(hist)
        return r

    @classmethod
    def philox(cls, ctr, key):
        c = np.asarray(ctr, np.uint32); k = np.asarray(key, np.uint32); o = np.zeros(4, np.uint32)
        cls.lib().oracle_philox4x32(_ptr(c), _ptr(k), _ptr(o))
        return o

    @classmethod
    def normal4(cls, seed, frame, block, row=0, stream=0):
        o = np.zeros(4, np.float32)
        cls.lib().oracle_normal4(seed, frame, block, row, stream, _ptr(o))
        return o


# ------------------------------------------------------------------------------------------------
class NbOracle:
    """Min-max decoding of non-binary codes, C restatement (oracle/ldpc_nb_oracle.c).  Parity unpinned."""

    def __init__(self, alist_path):
        L = Oracle.lib()
        L.oracle_nb_code_load_alist.restype = C.c_void_p
        L.oracle_nb_code_load_alist.argtypes = [C.c_char_p]
        L.oracle_nb_code_free.argtypes = [C.c_void_p]
        L.oracle_nb_dims.argtypes = [C.c_void_p] + [C.POINTER(C.c_int)] * 4
        L.oracle_nb_decode.argtypes = [C.c_void_p, C.c_int, C.c_int64, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]
        L.oracle_nb_simulate.argtypes = [C.c_void_p, C.c_int, C.c_double, C.c_double, C.c_uint64, C.c_int64, C.c_int64, C.c_void_p]
        self.L = L
        self.h = L.oracle_nb_code_load_alist(os.fsencode(alist_path))
        if not self.h:
            raise ValueError("oracle: cannot load " + alist_path)
        v = [C.c_int() for _ in range(4)]
        L.oracle_nb_dims(self.h, *[C.byref(x) for x in v])
        self.N, self.M, self.q, self.m = [x.value for x in v]

    def __del__(self):
        try:
            if getattr(self, "h", None):
                self.L.oracle_nb_code_free(self.h)
                self.h = None
        except Exception:
            pass

    @staticmethod
    def _counters(c):
        return {"errors": int(c[0]), "uncodedErrors": 0, "totalBits": int(c[1]), "totalWords": int(c[2]), "wordErrors": int(c[3]),
                "totalIterations": int(c[4]), "smoothingUsed": int(c[6]), "undetectedWords": int(c[5])}

    def decode(self, T, y):
        y = np.ascontiguousarray(y, np.float64)
        F = y.shape[0]
        sym, it, cnt = np.zeros((F, self.N), np.uint8), np.zeros(F, np.int32), np.zeros(8, np.int64)
        self.L.oracle_nb_decode(self.h, T, F, _ptr(y), _ptr(sym), _ptr(it), _ptr(cnt))
        return Result(symbols=sym, iters=it, counters=self._counters(cnt))

    def simulate(self, T, snr_db, R, seed, frame_begin, n_frames):
        cnt = np.zeros(8, np.int64)
        self.L.oracle_nb_simulate(self.h, T, snr_db, R, seed, frame_begin, n_frames, _ptr(cnt))
        return Result(counters=self._counters(cnt))


# ------------------------------------------------------------------------------------------------
class Reference:
    """One variant of the reference's own object code (oracle/_ref/libref_<variant>.so)."""

    _libs = {}

    @staticmethod
    def available(variant="decodeMinSum"):
        return os.path.exists(os.path.join(HERE, "_ref", "libref_%s.so" % variant))

    def __init__(self, variant, alist):
        self.variant = variant
        self.path = alist if os.path.exists(alist) else code_path(alist)
        if variant not in self._libs:
            L = C.CDLL(os.path.join(HERE, "_ref", "libref_%s.so" % variant))
            L.ref_decode_batch.argtypes = [C.c_char_p, C.POINTER(abi.DecoderCfg), C.POINTER(abi.Channel),
                                           C.POINTER(abi.Batch), C.POINTER(abi.Counters), C.c_void_p]
            L.ref_compiled_flags.restype = C.c_uint
            L.ref_run_main.argtypes = [C.c_int, C.POINTER(C.c_char_p), C.c_ulonglong, C.c_char_p]
            self._libs[variant] = L
        self.L = self._libs[variant]
        self.flags = int(self.L.ref_compiled_flags())
        self.kind = int(self.L.ref_kind())
        toks = open(self.path).read().split()[:2]
        self.N, self.M = int(toks[0]), int(toks[1])

    def decode(self, cfg, snr_db, R, y, noise=None, noise_rows=0, codeword=None, qpointer0=None, count=True):
        b, out = make_batch(self.N, y, noise, noise_rows, codeword, qpointer0)
        ch = abi.Channel(snr_db, R)
        cnt, hist = make_counters(self.N, cfg)
        qtrace = np.zeros(b.n_frames + 1, np.int32)
        rc = self.L.ref_decode_batch(self.path.encode(), C.byref(cfg), C.byref(ch), C.byref(b),
                                     C.byref(cnt) if count else None, _ptr(qtrace))
        if rc:
            raise RuntimeError("ref_decode_batch(%s) failed rc=%d" % (self.variant, rc))
        out["counters"] = cnt.as_dict()
        out.update(hist)
        out["d"] = unpack_bits(out.bits, self.N)
        out["qpointer_trace"] = qtrace
        return out

    def run_main(self, argv, stream_seed=1, stdout_path=None):
        return Reference.main(self.variant, argv, stream_seed, stdout_path)

    @staticmethod
    def main(variant, argv, stream_seed=1, stdout_path=None):
        """Run the variant's unmodified main() on the harness's deterministic random() stream.

        In a FRESH process: the reference keeps its parameters in mutable globals (num_iterations, theta, ...)
        which main() only partly re-initialises and which earlier ref_decode_batch calls in this process
        have overwritten."""
        import subprocess
        import sys
        code = ("import ctypes as C, os, sys\n"
                "L = C.CDLL(sys.argv[1])\n"
                "L.ref_run_main.argtypes = [C.c_int, C.POINTER(C.c_char_p), C.c_ulonglong, C.c_char_p]\n"
                "argv = sys.argv[4:]\n"
                "arr = (C.c_char_p * (len(argv) + 1))(*[a.encode() for a in argv], None)\n"
                "out = None if sys.argv[3] == '-' else os.fsencode(sys.argv[3])\n"
                "sys.exit(L.ref_run_main(len(argv), arr, int(sys.argv[2]), out) & 0xff)\n")
        lib = os.path.join(HERE, "_ref", "libref_%s.so" % variant)
        r = subprocess.run([sys.executable, "-c", code, lib, str(stream_seed), stdout_path or "-"] + list(argv))
        return r.returncode
