"""ctypes mirror of include/ldpc_gpu.h (structs and constants only; no library is loaded here)."""
import ctypes as C

OK = 0
ERR_INVALID_ARG, ERR_CUDA, ERR_NOMEM, ERR_UNSUPPORTED, ERR_COMM, ERR_IO, ERR_BAD_CODE = -1, -2, -3, -4, -5, -6, -7

KIND_MINSUM, KIND_BP, KIND_GDBF, KIND_NGDBF_HW, KIND_DDBMP, KIND_NGDBF_SC = 0, 1, 2, 3, 4, 5

F_QUANTIZE_SAMPLES = 1 << 0
F_SATURATE_SAMPLES = 1 << 1
F_NORMALIZED_MS = 1 << 2
F_OFFSET_MS = 1 << 3
F_SEQUENTIALMODE = 1 << 4
F_MODESWITCHING = 1 << 5
F_ADDNOISE = 1 << 6
F_WEIGHTSYNDROMES = 1 << 7
F_OUTPUTSMOOTHING = 1 << 8
F_THRESHOLDADAPTATION = 1 << 9
F_UNIFORMNOISE = 1 << 10
F_NOISESHAPING = 1 << 11
F_QUANTIZEPROBABILITIES = 1 << 12
F_REDECODE = 1 << 13
F_CERT_STOP = 1 << 14

# reference -D macro name -> flag bit (C_implementations/Makefile:24-71)
MACRO_FLAGS = {
    "quantizeSamples": F_QUANTIZE_SAMPLES, "saturateSamples": F_SATURATE_SAMPLES,
    "normalizedMS": F_NORMALIZED_MS, "offsetMS": F_OFFSET_MS,
    "sequentialmode": F_SEQUENTIALMODE, "modeswitching": F_MODESWITCHING, "addNoise": F_ADDNOISE,
    "weightSyndromes": F_WEIGHTSYNDROMES, "outputSmoothing": F_OUTPUTSMOOTHING,
    "thresholdAdaptation": F_THRESHOLDADAPTATION, "uniformNoise": F_UNIFORMNOISE,
    "noiseShaping": F_NOISESHAPING, "quantizeProbabilities": F_QUANTIZEPROBABILITIES, "redecode": F_REDECODE,
    "certStop": F_CERT_STOP,
}

PREC_F64, PREC_F32, PREC_F16X2 = 0, 1, 2
CHANNEL_EXACT, CHANNEL_FAST = 0, 1
MEM_HOST, MEM_DEVICE = 0, 1
DT_F64, DT_F32, DT_F16, DT_Q8, DT_QP = 0, 1, 2, 3, 4
HW_QBUF = 2648


class DecoderCfg(C.Structure):
    _fields_ = [
        ("kind", C.c_int32), ("flags", C.c_uint32), ("precision", C.c_int32), ("num_iterations", C.c_int32),
        ("Ymax", C.c_double), ("Q", C.c_int32), ("NQ", C.c_int32),
        ("alpha", C.c_double), ("delta", C.c_double), ("theta", C.c_double), ("lambda_", C.c_double),
        ("noiseScale", C.c_double), ("windowsize", C.c_int32), ("maxphase", C.c_int32),
        ("Tswitch", C.c_int32), ("channel_mode", C.c_int32),
        ("w", C.c_double), ("theta0", C.c_double), ("MAXLLR", C.c_double),
    ]


class Channel(C.Structure):
    _fields_ = [("snr_db", C.c_double), ("R", C.c_double)]


class Counters(C.Structure):
    _fields_ = [
        ("errors", C.c_int64), ("uncodedErrors", C.c_int64), ("totalBits", C.c_int64), ("totalWords", C.c_int64),
        ("wordErrors", C.c_int64), ("totalIterations", C.c_int64), ("smoothingUsed", C.c_int64),
        ("undetectedWords", C.c_int64),
        ("error_weight_hist", C.POINTER(C.c_int64)), ("iter_hist", C.POINTER(C.c_int64)),
        ("phase_hist", C.POINTER(C.c_int64)),
    ]
    SCALARS = ("errors", "uncodedErrors", "totalBits", "totalWords", "wordErrors", "totalIterations",
               "smoothingUsed", "undetectedWords")

    def as_dict(self):
        return {k: int(getattr(self, k)) for k in self.SCALARS}


class Batch(C.Structure):
    _fields_ = [
        ("n_frames", C.c_int64), ("mem", C.c_int32), ("y_dtype", C.c_int32),
        ("y", C.c_void_p), ("noise", C.c_void_p), ("noise_rows", C.c_int64),
        ("codeword", C.c_void_p), ("qpointer0", C.c_void_p),
        ("out_bits", C.c_void_p), ("out_iters", C.c_void_p), ("out_soft", C.c_void_p),
        ("out_errors", C.c_void_p), ("out_flags", C.c_void_p),
    ]


class SimArgs(C.Structure):
    _fields_ = [
        ("seed", C.c_uint64), ("frame_begin", C.c_int64), ("n_frames", C.c_int64),
        ("stop_errors", C.c_int64), ("stop_word_errors", C.c_int64), ("poll_frames", C.c_int64),
    ]


def default_cfg(kind, **kw):
    """Defaults = the reference's initial values of its parameter globals
    (src/decodeGDBF.cpp:48-56, src/NGDBFhw.cpp:48-57, src/decodeBP.cpp:58)."""
    c = DecoderCfg()
    c.kind = kind
    c.precision = PREC_F64
    c.num_iterations = 10
    c.Ymax, c.Q, c.NQ = 2.25, 5, 16
    c.alpha, c.delta, c.theta, c.lambda_ = 2.25, 0.0, -0.6, 0.991
    c.noiseScale, c.windowsize, c.maxphase, c.Tswitch = 1.0, 64, 7, 0
    c.w, c.theta0, c.MAXLLR = 0.185, -0.525, 20.0
    if kind == KIND_NGDBF_HW:
        c.num_iterations, c.Ymax, c.noiseScale, c.maxphase, c.NQ = 600, 1.625, 0.95, 1, 5
    if kind == KIND_MINSUM:
        c.alpha = 1.25
    if kind == KIND_NGDBF_SC:                    # SystemC/NGDBF/example.sh
        c.num_iterations, c.theta, c.lambda_, c.Q, c.Ymax, c.alpha, c.windowsize, c.noiseScale = 100, -0.5, 0.975, 4, 3.0, 0.95, 32, 1.0
    for k, v in kw.items():
        if k == "lambda":
            k = "lambda_"
        if k == "flags" and not isinstance(v, int):
            v = sum(MACRO_FLAGS[m] for m in v)
        if not hasattr(c, k):
            raise AttributeError(k)
        setattr(c, k, v)
    return c


def gdbf_rows_per_step(flags):
    return (1 if flags & F_ADDNOISE else 0) + (1 if flags & F_QUANTIZEPROBABILITIES else 0)


def sc_noise_len(cfg, N):
    """Entries per frame of batch.noise for KIND_NGDBF_SC: the per-frame window of the node-to-node noise chain."""
    return N + cfg.num_iterations + 1


def noise_rows_needed(cfg):
    if cfg.kind != KIND_GDBF:
        return 0
    ph = max(1, cfg.maxphase) if cfg.flags & F_REDECODE else 1
    return cfg.num_iterations * ph * gdbf_rows_per_step(cfg.flags)


def iter_hist_len(cfg):
    ph = cfg.maxphase if (cfg.flags & F_REDECODE and cfg.kind == KIND_GDBF and cfg.maxphase > 1) else 1
    return cfg.num_iterations * ph + 1


def quantizer_levels(y, Ymax, Q):
    """LDPC_GPU_DT_Q8 encoding of raw samples: what quantize() (src/decodeMinSum.cpp:480-489) makes of y, as signed levels."""
    import numpy as np
    y = np.asarray(y, np.float64)
    Nq = 2.0 ** Q                      # the reference passes Nq = 2^Q and quantises to (Nq - 1) intervals (src/decodeMinSum.cpp:133,480-489)
    a = np.abs(y)
    L = np.floor(a * (Nq - 1.0) / (2.0 * Ymax))
    k = np.where(a > Ymax, 32.0, np.maximum(1.0, L))
    return np.where(y >= 0, k, -k).astype(np.int8)


def quantizer_levels_packed(y, Ymax, Q):
    """LDPC_GPU_DT_QP encoding: quantizer_levels() packed Q bits per sample, [F][N*Q/8] bytes."""
    import numpy as np
    k = quantizer_levels(y, Ymax, Q).astype(np.int32)
    F, N = k.shape
    assert (N * Q) % 32 == 0
    mag = np.minimum(np.abs(k), 2 ** (Q - 1)) - 1
    code = (mag | ((k < 0).astype(np.int32) << (Q - 1))).astype(np.uint8)
    bits = ((code[:, :, None] >> np.arange(Q)) & 1).astype(np.uint8).reshape(F, N * Q)
    return np.packbits(bits, axis=1, bitorder="little")
