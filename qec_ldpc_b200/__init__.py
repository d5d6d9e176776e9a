"""qec_ldpc_b200 -- B200-native Monte-Carlo BP decoder for quasi-cyclic quantum CSS LDPC codes.

The product is the C-ABI shared library ``lib/libqldpc_b200.so`` (``include/qldpc_b200.h``) and the C++ classes in
``cpp/`` that mirror the reference's ``Decoder`` / ``DecoderGPU`` / ``Quantum_LDPC_Code`` / ``CodeStatistics``.
This module is only the ctypes plumbing the tests and ``bench.py`` use to call that ABI from Python; it contains
no decoding logic and no CPU fallback: if the library is missing it raises, and on a machine without a CUDA
device ``Decoder(...)`` raises ``QldpcError`` (QLDPC_ERR_NO_DEVICE).
"""
import ctypes as C
import os

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(HERE, "lib", "libqldpc_b200.so")

NUM_COUNTERS = 12
COUNTER_NAMES = ["frames", "xTested", "zTested", "corrected", "synX", "synZ", "logical", "cvX", "cvZ", "itersX",
                 "itersZ", "nanFrames"]
SYNDROME_FAIL_X, SYNDROME_FAIL_Z, CONVERGENCE_FAIL_X, CONVERGENCE_FAIL_Z = 1, 2, 4, 8
FRAME_LOGICAL, FRAME_CORRECTED, FRAME_NAN = 16, 32, 64
ERR_ARG, ERR_CUDA, ERR_NO_DEVICE, ERR_UNSUPPORTED, ERR_IO = -1, -2, -3, -4, -5


class QldpcError(RuntimeError):
    def __init__(self, code, msg):
        super().__init__("qldpc error %d: %s" % (code, msg))
        self.code = code


class CodeInfo(C.Structure):
    _fields_ = [(k, C.c_int32) for k in
                ["J", "K", "L", "P", "sigma", "tau", "n", "mX", "mZ", "dcX", "dcZ", "dvX", "dvZ", "EX", "EZ",
                 "logical_rows", "is_qc", "logical_from_file"]]


_lib = None


def load_library():
    """Loads the CUDA-backed C-ABI library; raises if it has not been built (no fallback)."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise ImportError("%s not built: run `python -m qec_ldpc_b200.build` (needs nvcc); there is no fallback path"
                          % LIB_PATH)
    L = C.CDLL(LIB_PATH)
    vp, i32, i64, u32, u64, f32 = C.c_void_p, C.c_int, C.c_int64, C.c_uint32, C.c_uint64, C.c_float
    L.qldpc_version.restype = C.c_char_p
    L.qldpc_last_error.restype = C.c_char_p
    L.qldpc_code_create_qc.argtypes = [i32] * 6 + [C.POINTER(vp)]
    L.qldpc_code_create_dense.argtypes = [i32] * 6 + [vp, vp, vp, C.POINTER(vp)]
    L.qldpc_code_create_from_file.argtypes = [C.c_char_p, C.POINTER(vp)]
    L.qldpc_code_write_file.argtypes = [vp, C.c_char_p]
    L.qldpc_code_destroy.argtypes = [vp]
    L.qldpc_code_destroy.restype = None
    L.qldpc_code_get_info.argtypes = [vp, C.POINTER(CodeInfo)]
    L.qldpc_code_name.argtypes = [vp, C.c_char_p, i32]
    L.qldpc_code_exponents.argtypes = [vp, i32, vp]
    L.qldpc_code_csr.argtypes = [vp, i32, vp]
    L.qldpc_code_csc.argtypes = [vp, i32, vp, vp]
    L.qldpc_code_dense.argtypes = [vp, i32, vp]
    L.qldpc_code_is_css.argtypes = [vp]
    L.qldpc_code_syndrome.argtypes = [vp, i32, vp, vp]
    L.qldpc_code_check_logical.argtypes = [vp, vp]
    L.qldpc_decoder_create.argtypes = [vp, i32, i32, C.POINTER(vp)]
    L.qldpc_decoder_destroy.argtypes = [vp]
    L.qldpc_decoder_destroy.restype = None
    L.qldpc_decoder_set_stream.argtypes = [vp, vp]
    L.qldpc_decoder_configure.argtypes = [vp, i32, i32, i32, i32]
    L.qldpc_decoder_set_host_threads.argtypes = [vp, i32]
    L.qldpc_decoder_launch_info.argtypes = [vp, i32, vp]
    L.qldpc_decoder_host_threads_in_use.argtypes = [vp, i32]
    L.qldpc_decode_batch.argtypes = [vp, vp, vp, i64, f32, i32, vp, vp, vp, vp]
    L.qldpc_decode_batch_device.argtypes = [vp, vp, vp, i64, f32, i32, vp, vp, vp, vp]
    L.qldpc_get_statistics_weightw.argtypes = [vp, i32, i64, f32, i32, u32, vp, vp, vp]
    L.qldpc_get_statistics_depolarizing.argtypes = [vp, u64, u64, i64, f32, i32, vp, vp, vp]
    L.qldpc_get_stats_from_errors_i32.argtypes = [vp, vp, vp, i64, f32, i32, vp, vp, vp]
    L.qldpc_get_stats_from_errors_u8.argtypes = [vp, vp, vp, i64, f32, i32, vp, vp, vp]
    L.qldpc_decoder_enable_timing.argtypes = [vp, i32]
    L.qldpc_decoder_get_timing.argtypes = [vp, vp, vp, i32]
    L.qldpc_debug_division_check.argtypes = [vp, u64, i64, vp]
    L.qldpc_debug_weightw_patterns.argtypes = [u32, i32, i32, i64, i32, vp, vp]
    L.qldpc_debug_host_pack.argtypes = [vp, i32, i64, i32, vp, i32]
    L.qldpc_debug_host_unpack.argtypes = [vp, i64, i32, vp, i32]
    L.qldpc_debug_host_read_gbs.argtypes = [vp, i64, i32, i32, C.POINTER(C.c_double)]
    L.qldpc_debug_generate.argtypes = [vp, u64, u64, i64, f32, vp, vp, vp, vp]
    L.qldpc_debug_bp_trace.argtypes = [vp, i32, vp, i32, f32, i32, i32, vp, vp, vp]
    _lib = L
    return L


def _check(rc):
    if rc < 0:
        raise QldpcError(rc, load_library().qldpc_last_error().decode())
    return rc


def _ptr(a):
    return None if a is None else a.ctypes.data_as(C.c_void_p)


def host_pack(rows, threads=4):
    """Test tap: bit-packs a 2-D uint8 / int32 array row by row with the library's host packer."""
    a = np.ascontiguousarray(rows)
    assert a.ndim == 2 and a.dtype in (np.uint8, np.int32)
    out = np.zeros((a.shape[0], (a.shape[1] + 31) // 32), np.uint32)
    _check(load_library().qldpc_debug_host_pack(_ptr(a), a.dtype.itemsize, a.shape[0], a.shape[1], _ptr(out), threads))
    return out


def weightw_patterns(seed, weight, n, nframes, threads=4):
    """Test tap: the packed x / z rows of the reference-compatible weight-W error stream (no GPU needed)."""
    nw = (n + 31) // 32
    x, z = np.zeros((nframes, nw), np.uint32), np.zeros((nframes, nw), np.uint32)
    _check(load_library().qldpc_debug_weightw_patterns(seed, weight, n, nframes, threads, _ptr(x), _ptr(z)))
    return x, z


def host_read_gbs(ptr, nbytes, threads, repeats=3):
    """Measured streaming-read bandwidth (GB/s) of a host buffer with the packer's worker threads."""
    out = C.c_double(0.0)
    _check(load_library().qldpc_debug_host_read_gbs(C.c_void_p(ptr), nbytes, threads, repeats, C.byref(out)))
    return out.value


def default_host_threads():
    """Worker threads the host-buffer entry points use by default on this process (host_pack.h)."""
    return int(load_library().qldpc_default_host_threads())


def host_unpack(words, cols, threads=4):
    w = np.ascontiguousarray(words, np.uint32)
    out = np.zeros((w.shape[0], cols), np.uint8)
    _check(load_library().qldpc_debug_host_unpack(_ptr(w), w.shape[0], cols, _ptr(out), threads))
    return out


class Code:
    """Handle on a ``qldpc_code`` (Quantum_LDPC_Code + packed CSR/CSC edge tables)."""

    def __init__(self, handle):
        self._lib = load_library()
        self.h = handle
        info = CodeInfo()
        _check(self._lib.qldpc_code_get_info(self.h, C.byref(info)))
        self.info = info
        for k, _ in CodeInfo._fields_:
            setattr(self, k, int(getattr(info, k)))
        self.m = (self.mX, self.mZ)
        self.dc = (self.dcX, self.dcZ)
        self.dv = (self.dvX, self.dvZ)
        self.E = (self.EX, self.EZ)

    @classmethod
    def qc(cls, J, K, L, P, sigma, tau):
        lib = load_library()
        h = C.c_void_p()
        _check(lib.qldpc_code_create_qc(J, K, L, P, sigma, tau, C.byref(h)))
        return cls(h)

    @classmethod
    def dense(cls, J, K, L, P, sigma, tau, pcmX, pcmZ, iMinusP=None):
        lib = load_library()
        h = C.c_void_p()
        x = np.ascontiguousarray(pcmX, np.int32)
        z = np.ascontiguousarray(pcmZ, np.int32)
        i = None if iMinusP is None else np.ascontiguousarray(iMinusP, np.int32)
        _check(lib.qldpc_code_create_dense(J, K, L, P, sigma, tau, _ptr(x), _ptr(z), _ptr(i), C.byref(h)))
        return cls(h)

    @classmethod
    def from_file(cls, path):
        lib = load_library()
        h = C.c_void_p()
        _check(lib.qldpc_code_create_from_file(os.fsencode(path), C.byref(h)))
        return cls(h)

    def __del__(self):
        try:
            if self.h:
                self._lib.qldpc_code_destroy(self.h)
                self.h = None
        except Exception:
            pass

    def write_file(self, path):
        _check(self._lib.qldpc_code_write_file(self.h, os.fsencode(path)))

    def name(self):
        buf = C.create_string_buffer(256)
        _check(self._lib.qldpc_code_name(self.h, buf, 256))
        return buf.value.decode()

    def exponents(self, side):
        rows = self.J if side == 0 else self.K
        out = np.zeros((rows, self.info.L), np.int32)
        _check(self._lib.qldpc_code_exponents(self.h, side, _ptr(out)))
        return out

    def csr(self, side):
        out = np.zeros((self.m[side], self.dc[side]), np.int32)
        _check(self._lib.qldpc_code_csr(self.h, side, _ptr(out)))
        return out

    def csc(self, side):
        a = np.zeros((self.n, self.dv[side]), np.int32)
        b = np.zeros((self.n, self.dv[side]), np.int32)
        _check(self._lib.qldpc_code_csc(self.h, side, _ptr(a), _ptr(b)))
        return a, b

    def dense_matrix(self, which):
        shape = [(self.mX, self.n), (self.mZ, self.n), (self.logical_rows, 2 * self.n), (2 * self.n, 2 * self.n)][which]
        out = np.zeros(shape, np.int32)
        _check(self._lib.qldpc_code_dense(self.h, which, _ptr(out)))
        return out

    def is_css(self):
        return bool(_check(self._lib.qldpc_code_is_css(self.h)))

    def syndrome(self, side, err):
        e = np.ascontiguousarray(err, np.int32)
        s = np.zeros(self.m[side], np.int32)
        _check(self._lib.qldpc_code_syndrome(self.h, side, _ptr(e), _ptr(s)))
        return s

    def check_logical(self, err2n):
        e = np.ascontiguousarray(err2n, np.int32)
        return bool(_check(self._lib.qldpc_code_check_logical(self.h, _ptr(e))))


class Decoder:
    """Handle on a ``qldpc_decoder`` (the DecoderGPU state on one device)."""

    def __init__(self, code, device=-1, max_frames=0):
        self._lib = load_library()
        self.code = code
        self.h = C.c_void_p()
        _check(self._lib.qldpc_decoder_create(code.h, device, max_frames, C.byref(self.h)))

    def __del__(self):
        try:
            if self.h:
                self._lib.qldpc_decoder_destroy(self.h)
                self.h = None
        except Exception:
            pass

    def set_stream(self, stream_ptr):
        _check(self._lib.qldpc_decoder_set_stream(self.h, C.c_void_p(stream_ptr)))

    def configure(self, side, frames_per_tile=0, threads=0, ctas_per_sm=0):
        """Launch shape of one side (0 = heuristic).  frames_per_tile=-1 forces the HBM-resident path; `threads` is
        then the number of frame slots kept in flight."""
        _check(self._lib.qldpc_decoder_configure(self.h, side, frames_per_tile, threads, ctas_per_sm))

    def set_host_threads(self, threads):
        """Worker threads that pack host rows before the H2D copy (-1 default, 0 = copy raw rows, pack on device)."""
        _check(self._lib.qldpc_decoder_set_host_threads(self.h, threads))

    def host_threads_in_use(self, elem=4):
        """Threads the host-buffer entry points use for rows of `elem`-byte elements (0 = raw rows over the link)."""
        return _check(self._lib.qldpc_decoder_host_threads_in_use(self.h, elem))

    def launch_info(self, side):
        out = np.zeros(8, np.int32)
        _check(self._lib.qldpc_decoder_launch_info(self.h, side, _ptr(out)))
        keys = ["vec", "threads", "ctas_per_sm", "grid", "smem", "regs", "num_sms", "chunk"]
        return dict(zip(keys, [int(v) for v in out]))

    TIMER_NAMES = ["generate", "syndrome", "bp_x", "bp_z", "stats", "pack"]

    def enable_timing(self, on=True):
        _check(self._lib.qldpc_decoder_enable_timing(self.h, int(on)))

    def get_timing(self, reset=True):
        ms = np.zeros(6, np.float64)
        n = np.zeros(6, np.uint64)
        _check(self._lib.qldpc_decoder_get_timing(self.h, _ptr(ms), _ptr(n), int(reset)))
        return dict(zip(self.TIMER_NAMES, ms.tolist())), dict(zip(self.TIMER_NAMES, [int(v) for v in n]))

    def decode_batch(self, synX, synZ, p, maxit, want_iters=True):
        sx = np.ascontiguousarray(synX, np.uint8)
        sz = np.ascontiguousarray(synZ, np.uint8)
        nf, n = sx.shape[0], self.code.n
        ox, oz = np.zeros((nf, n), np.uint8), np.zeros((nf, n), np.uint8)
        fl = np.zeros(nf, np.uint8)
        it = np.zeros((nf, 2), np.uint32) if want_iters else None
        _check(self._lib.qldpc_decode_batch(self.h, _ptr(sx), _ptr(sz), nf, p, maxit, _ptr(ox), _ptr(oz), _ptr(fl), _ptr(it)))
        return ox, oz, fl, it

    def decode_batch_ptr(self, synX_ptr, synZ_ptr, nframes, p, maxit, outX_ptr, outZ_ptr, flags_ptr, iters_ptr=0):
        """Host pointers (e.g. pinned torch tensors): bytes in, bytes out, as qldpc_decode_batch documents."""
        _check(self._lib.qldpc_decode_batch(self.h, C.c_void_p(synX_ptr), C.c_void_p(synZ_ptr), nframes, p, maxit,
                                            C.c_void_p(outX_ptr), C.c_void_p(outZ_ptr), C.c_void_p(flags_ptr),
                                            C.c_void_p(iters_ptr) if iters_ptr else None))

    def decode_batch_device(self, d_synX, d_synZ, nframes, p, maxit, d_outX, d_outZ, d_flags, d_iters=0):
        _check(self._lib.qldpc_decode_batch_device(self.h, C.c_void_p(d_synX), C.c_void_p(d_synZ), nframes, p, maxit,
                                                C.c_void_p(d_outX), C.c_void_p(d_outZ), C.c_void_p(d_flags),
                                                C.c_void_p(d_iters) if d_iters else None))

    def _stats(self, fn, args, nframes, per_frame):
        k = np.zeros(NUM_COUNTERS, np.uint64)
        fl = np.zeros(nframes, np.uint8) if per_frame else None
        it = np.zeros((nframes, 2), np.uint32) if per_frame else None
        _check(fn(self.h, *args, _ptr(k), _ptr(fl), _ptr(it)))
        return dict(counters=k, flags=fl, iters=it)

    def get_statistics_weightw(self, W, count, p, maxit, seed, per_frame=False):
        return self._stats(self._lib.qldpc_get_statistics_weightw, (W, count, p, maxit, seed), count, per_frame)

    def get_statistics_depolarizing(self, seed, first_frame, nframes, p, maxit, per_frame=False):
        return self._stats(self._lib.qldpc_get_statistics_depolarizing, (seed, first_frame, nframes, p, maxit), nframes,
                           per_frame)

    def get_stats_from_errors(self, xerr, zerr, p, maxit, per_frame=False):
        if xerr.dtype == np.int32:
            x, z = np.ascontiguousarray(xerr, np.int32), np.ascontiguousarray(zerr, np.int32)
            fn = self._lib.qldpc_get_stats_from_errors_i32
        else:
            x, z = np.ascontiguousarray(xerr, np.uint8), np.ascontiguousarray(zerr, np.uint8)
            fn = self._lib.qldpc_get_stats_from_errors_u8
        return self._stats(fn, (_ptr(x), _ptr(z), x.shape[0], p, maxit), x.shape[0], per_frame)

    def get_stats_from_errors_ptr(self, x_ptr, z_ptr, nframes, p, maxit, elem=4):
        """Host pointers (e.g. pinned torch tensors); elem = 4 (int32, the reference's layout) or 1 (bytes)."""
        fn = self._lib.qldpc_get_stats_from_errors_i32 if elem == 4 else self._lib.qldpc_get_stats_from_errors_u8
        k = np.zeros(NUM_COUNTERS, np.uint64)
        _check(fn(self.h, C.c_void_p(x_ptr), C.c_void_p(z_ptr), nframes, p, maxit, _ptr(k), None, None))
        return k

    def debug_division_check(self, seed, npairs):
        out = np.zeros(3, np.uint64)
        _check(self._lib.qldpc_debug_division_check(self.h, seed, npairs, _ptr(out)))
        return dict(mismatches=int(out[0]), deferred=int(out[1]), zero_numerators=int(out[2]))

    def debug_generate(self, seed, first_frame, nframes, p):
        n, mX, mZ = self.code.n, self.code.mX, self.code.mZ
        x, z = np.zeros((nframes, n), np.uint8), np.zeros((nframes, n), np.uint8)
        sx, sz = np.zeros((nframes, mX), np.uint8), np.zeros((nframes, mZ), np.uint8)
        _check(self._lib.qldpc_debug_generate(self.h, seed, first_frame, nframes, p, _ptr(x), _ptr(z), _ptr(sx), _ptr(sz)))
        return x, z, sx, sz

    def debug_bp_trace(self, side, syn, p, maxit, cap):
        s = np.ascontiguousarray(syn, np.uint8)
        nf, E = s.shape[0], self.code.E[side]
        q = np.zeros((nf, cap, E), np.float32)
        r = np.zeros((nf, cap, E), np.float32)
        it = np.zeros(nf, np.uint32)
        _check(self._lib.qldpc_debug_bp_trace(self.h, side, _ptr(s), nf, p, maxit, cap, _ptr(q), _ptr(r), _ptr(it)))
        return q, r, it
