"""TEST INFRASTRUCTURE ONLY -- ctypes bindings for the CPU checkers under oracle/.

* ``Oracle``    : oracle/liboracle.so, the plain-C restatement (oracle.c).
* ``Reference`` : oracle/_ref/libqldpc_ref.so, the UNMODIFIED reference compiled from /root/reference
                  (oracle/ref_harness.cpp); present when built in the authoring container (it travels to the
                  GPU box as a prebuilt file), absent otherwise.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / ``--impl reference`` legs import this module.
The product package (qec_ldpc_b200) never does.
"""
import ctypes as C
import os
import subprocess
import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
REF_ROOT = "/root/reference"

u8p = np.ctypeslib.ndpointer(np.uint8, flags="C_CONTIGUOUS")
i32p = np.ctypeslib.ndpointer(np.int32, flags="C_CONTIGUOUS")
f32p = np.ctypeslib.ndpointer(np.float32, flags="C_CONTIGUOUS")
u64p = np.ctypeslib.ndpointer(np.uint64, flags="C_CONTIGUOUS")
u32p = np.ctypeslib.ndpointer(np.uint32, flags="C_CONTIGUOUS")

COUNTER_NAMES = ["frames", "xTested", "zTested", "corrected", "synX", "synZ", "logical", "cvX", "cvZ", "itersX",
                 "itersZ", "nanFrames"]


def build(ref=True):
    """Compile the checkers (not the product).  gcc/g++ from PATH; see oracle/Makefile."""
    env = dict(os.environ)
    env.pop("CC", None)
    env.pop("CXX", None)
    subprocess.run(["make", "-s", "-C", HERE, "liboracle.so"], check=True, env=env)
    if ref and os.path.isdir(os.path.join(REF_ROOT, "QEC_LDPC")):
        subprocess.run(["make", "-s", "-C", HERE, "ref"], check=True, env=env, stdout=subprocess.DEVNULL)


def _opt(a, ptr_t=C.c_void_p):
    return None if a is None else a.ctypes.data_as(ptr_t)


class Oracle:
    def __init__(self):
        path = os.path.join(HERE, "liboracle.so")
        if not os.path.exists(path):
            build(ref=False)
        L = self.L = C.CDLL(path)
        L.oracle_code_qc.restype = C.c_void_p
        L.oracle_code_qc.argtypes = [C.c_int] * 6
        L.oracle_code_dense.restype = C.c_void_p
        L.oracle_code_dense.argtypes = [C.c_int] * 6 + [i32p, i32p]
        L.oracle_code_free.argtypes = [C.c_void_p]
        L.oracle_code_info.argtypes = [C.c_void_p, i32p]
        L.oracle_code_tables.argtypes = [C.c_void_p, C.c_int, i32p, i32p, i32p]
        L.oracle_qc_exponents.argtypes = [C.c_int] * 6 + [i32p, i32p]
        L.oracle_dense_pcm.argtypes = [C.c_void_p, C.c_int, i32p]
        L.oracle_set_logical.argtypes = [C.c_void_p, i32p, C.c_int]
        L.oracle_philox4x32_10.argtypes = [u32p, u32p, u32p]
        L.oracle_depolarizing_thresholds.argtypes = [C.c_float, u32p]
        L.oracle_depolarizing.argtypes = [C.c_void_p, C.c_uint64, C.c_uint64, C.c_float, u8p, u8p]
        L.oracle_depolarizing_bulk.argtypes = [C.c_void_p, C.c_uint64, C.c_uint64, C.c_int, C.c_float, u8p, u8p]
        L.oracle_syndrome.argtypes = [C.c_void_p, C.c_int, u8p, u8p]
        L.oracle_bp.restype = C.c_int
        L.oracle_bp.argtypes = [C.c_void_p, C.c_int, u8p, C.c_float, C.c_int, f32p, f32p, C.c_void_p, C.c_void_p, C.c_int]
        L.oracle_decode.restype = C.c_int
        L.oracle_decode.argtypes = [C.c_void_p, u8p, u8p, C.c_float, C.c_int, u8p, u8p, i32p]
        L.oracle_check_logical.restype = C.c_int
        L.oracle_check_logical.argtypes = [C.c_void_p, u8p]
        L.oracle_weightw_stream.argtypes = [C.c_uint32, C.c_int, C.c_int, C.c_int, u8p, u8p]
        L.oracle_run_frames.restype = C.c_double
        L.oracle_run_frames.argtypes = [C.c_void_p, u8p, u8p, C.c_int, C.c_float, C.c_int, C.c_int, u64p, C.c_void_p,
                                        C.c_void_p, C.c_void_p, C.c_void_p]
        L.oracle_run_depolarizing.restype = C.c_double
        L.oracle_run_depolarizing.argtypes = [C.c_void_p, C.c_uint64, C.c_uint64, C.c_int, C.c_float, C.c_int, C.c_int,
                                              u64p, C.c_void_p, C.c_void_p]
        L.oracle_get_statistics_weightw.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_float, C.c_int, C.c_uint32,
                                                    C.c_int, u64p]
        L.oracle_max_threads.restype = C.c_int

    def max_threads(self):
        return self.L.oracle_max_threads()

    def code_qc(self, J, K, L, P, sigma, tau):
        h = self.L.oracle_code_qc(J, K, L, P, sigma, tau)
        if not h:
            raise ValueError("irregular or invalid QC parameters")
        return OracleCode(self, h)

    def code_dense(self, J, K, L, P, sigma, tau, pcmX, pcmZ):
        h = self.L.oracle_code_dense(J, K, L, P, sigma, tau, np.ascontiguousarray(pcmX, np.int32).ravel(),
                                     np.ascontiguousarray(pcmZ, np.int32).ravel())
        if not h:
            raise ValueError("parity-check matrices are not regular")
        return OracleCode(self, h)

    def exponents(self, J, K, L, P, sigma, tau):
        hc = np.zeros(J * L, np.int32)
        hd = np.zeros(K * L, np.int32)
        self.L.oracle_qc_exponents(J, K, L, P, sigma, tau, hc, hd)
        return hc.reshape(J, L), hd.reshape(K, L)

    def philox(self, ctr, key):
        out = np.zeros(4, np.uint32)
        self.L.oracle_philox4x32_10(np.asarray(ctr, np.uint32), np.asarray(key, np.uint32), out)
        return out

    def thresholds(self, p):
        t = np.zeros(3, np.uint32)
        self.L.oracle_depolarizing_thresholds(p, t)
        return t

    def weightw_stream(self, seed, W, n, nframes):
        x = np.zeros((nframes, n), np.uint8)
        z = np.zeros((nframes, n), np.uint8)
        self.L.oracle_weightw_stream(seed, W, n, nframes, x, z)
        return x, z


class OracleCode:
    def __init__(self, lib, h):
        self.lib, self.L, self.h = lib, lib.L, h
        info = np.zeros(16, np.int32)
        self.L.oracle_code_info(h, info)
        (self.J, self.K, self.Lc, self.P, self.sigma, self.tau, self.n, mX, mZ, dcX, dcZ, dvX, dvZ, EX, EZ,
         _) = [int(v) for v in info]
        self.m, self.dc, self.dv, self.E = (mX, mZ), (dcX, dcZ), (dvX, dvZ), (EX, EZ)

    def __del__(self):
        try:
            self.L.oracle_code_free(self.h)
        except Exception:
            pass

    def tables(self, side):
        E = self.E[side]
        a, b, c = np.zeros(E, np.int32), np.zeros(E, np.int32), np.zeros(E, np.int32)
        self.L.oracle_code_tables(self.h, side, a, b, c)
        return a.reshape(self.m[side], self.dc[side]), b.reshape(self.n, self.dv[side]), c.reshape(self.n, self.dv[side])

    def dense(self, side):
        out = np.zeros(self.m[side] * self.n, np.int32)
        self.L.oracle_dense_pcm(self.h, side, out)
        return out.reshape(self.m[side], self.n)

    def set_logical(self, mat):
        mat = np.ascontiguousarray(mat, np.int32)
        assert mat.shape[1] == 2 * self.n
        self.L.oracle_set_logical(self.h, mat.ravel(), mat.shape[0])

    def depolarizing(self, seed, frame, p):
        x, z = np.zeros(self.n, np.uint8), np.zeros(self.n, np.uint8)
        self.L.oracle_depolarizing(self.h, seed, frame, p, x, z)
        return x, z

    def depolarizing_bulk(self, seed, first_frame, nframes, p):
        x, z = np.zeros((nframes, self.n), np.uint8), np.zeros((nframes, self.n), np.uint8)
        self.L.oracle_depolarizing_bulk(self.h, seed, first_frame, nframes, p, x, z)
        return x, z

    def syndrome(self, side, err):
        s = np.zeros(self.m[side], np.uint8)
        self.L.oracle_syndrome(self.h, side, np.ascontiguousarray(err, np.uint8), s)
        return s

    def bp(self, side, syn, p, maxit, trace=0):
        E = self.E[side]
        q, r = np.zeros(E, np.float32), np.zeros(E, np.float32)
        qt = np.zeros((trace, E), np.float32) if trace else None
        rt = np.zeros((trace, E), np.float32) if trace else None
        it = self.L.oracle_bp(self.h, side, np.ascontiguousarray(syn, np.uint8), p, maxit, q, r, _opt(qt), _opt(rt), trace)
        return it, q, r, qt, rt

    def decode(self, synX, synZ, p, maxit):
        ox, oz = np.zeros(self.n, np.uint8), np.zeros(self.n, np.uint8)
        iters = np.zeros(2, np.int32)
        code = self.L.oracle_decode(self.h, np.ascontiguousarray(synX, np.uint8), np.ascontiguousarray(synZ, np.uint8),
                                    p, maxit, ox, oz, iters)
        return code, ox, oz, iters

    def check_logical(self, err2n):
        return self.L.oracle_check_logical(self.h, np.ascontiguousarray(err2n, np.uint8))

    def run_frames(self, xerr, zerr, p, maxit, nthreads=0, want_out=False):
        nf = xerr.shape[0]
        k = np.zeros(12, np.uint64)
        flags = np.zeros(nf, np.uint8)
        iters = np.zeros((nf, 2), np.uint8)
        ox = np.zeros((nf, self.n), np.uint8) if want_out else None
        oz = np.zeros((nf, self.n), np.uint8) if want_out else None
        sec = self.L.oracle_run_frames(self.h, np.ascontiguousarray(xerr, np.uint8), np.ascontiguousarray(zerr, np.uint8),
                                       nf, p, maxit, nthreads, k, _opt(flags), _opt(iters), _opt(ox), _opt(oz))
        return dict(counters=k, flags=flags, iters=iters, outX=ox, outZ=oz, seconds=sec)

    def run_depolarizing(self, seed, first_frame, nframes, p, maxit, nthreads=0):
        k = np.zeros(12, np.uint64)
        flags = np.zeros(nframes, np.uint8)
        iters = np.zeros((nframes, 2), np.uint8)
        sec = self.L.oracle_run_depolarizing(self.h, seed, first_frame, nframes, p, maxit, nthreads, k, _opt(flags),
                                             _opt(iters))
        return dict(counters=k, flags=flags, iters=iters, seconds=sec)

    def get_statistics_weightw(self, W, count, p, maxit, seed, nthreads=0):
        k = np.zeros(12, np.uint64)
        self.L.oracle_get_statistics_weightw(self.h, W, count, p, maxit, seed, nthreads, k)
        return k


class Reference:
    """The unmodified reference DecoderCPU behind oracle/ref_harness.cpp."""

    @staticmethod
    def available():
        return os.path.exists(os.path.join(HERE, "_ref", "libqldpc_ref.so"))

    def __init__(self):
        L = self.L = C.CDLL(os.path.join(HERE, "_ref", "libqldpc_ref.so"))
        L.qref_code_from_file.restype = C.c_void_p
        L.qref_code_from_file.argtypes = [C.c_char_p]
        L.qref_code_free.argtypes = [C.c_void_p]
        L.qref_code_dims.argtypes = [C.c_void_p, i32p]
        L.qref_code_dense.argtypes = [C.c_void_p, C.c_int, i32p]
        L.qref_code_name.argtypes = [C.c_void_p, C.c_char_p, C.c_int]
        L.qref_decoder_create.restype = C.c_void_p
        L.qref_decoder_create.argtypes = [C.c_void_p]
        L.qref_decoder_free.argtypes = [C.c_void_p]
        L.qref_get_statistics.restype = C.c_longlong
        L.qref_get_statistics.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_float, C.c_int, C.c_uint, C.c_int, u32p]
        L.qref_format_statistics.restype = C.c_int
        L.qref_format_statistics.argtypes = [C.c_void_p, u32p, C.c_uint, C.c_longlong, C.c_char_p, C.c_int]
        L.qref_decode.restype = C.c_int
        L.qref_decode.argtypes = [C.c_void_p, i32p, i32p, C.c_float, C.c_int, i32p, i32p]
        L.qref_syndrome.argtypes = [C.c_void_p, C.c_int, i32p, i32p]
        L.qref_check_logical.restype = C.c_int
        L.qref_check_logical.argtypes = [C.c_void_p, i32p]
        L.qref_bp_trace.restype = C.c_int
        L.qref_bp_trace.argtypes = [C.c_void_p, C.c_int, i32p, C.c_float, C.c_int, C.c_int, f32p, f32p, i32p]
        L.qref_run_frames.restype = C.c_double
        L.qref_run_frames.argtypes = [C.c_void_p, u8p, u8p, C.c_int, C.c_float, C.c_int, C.c_int, u64p, C.c_void_p,
                                      C.c_void_p, C.c_void_p]
        L.qref_max_threads.restype = C.c_int

    def max_threads(self):
        return self.L.qref_max_threads()

    def code_from_file(self, path):
        h = self.L.qref_code_from_file(path.encode())
        if not h:
            raise FileNotFoundError(path)
        return ReferenceCode(self, h)


class ReferenceCode:
    def __init__(self, lib, h):
        self.L, self.h = lib.L, h
        d = np.zeros(9, np.int32)
        self.L.qref_code_dims(h, d)
        self.J, self.K, self.Lc, self.P, self.sigma, self.tau, self.n, self.mX, self.mZ = [int(v) for v in d]
        self.m = (self.mX, self.mZ)
        self.dec = self.L.qref_decoder_create(h)

    def dense(self, which):
        shape = [(self.mX, self.n), (self.mZ, self.n), (2 * self.n, 2 * self.n)][which]
        out = np.zeros(shape[0] * shape[1], np.int32)
        self.L.qref_code_dense(self.h, which, out)
        return out.reshape(shape)

    def name(self):
        buf = C.create_string_buffer(256)
        self.L.qref_code_name(self.h, buf, 256)
        return buf.value.decode()

    def get_statistics(self, W, count, p, maxit, seed, nthreads=0):
        out = np.zeros(10, np.uint32)
        dur = self.L.qref_get_statistics(self.dec, W, count, p, maxit, seed, nthreads, out)
        return out, dur

    def format_statistics(self, v10, seed, dur):
        buf = C.create_string_buffer(4096)
        self.L.qref_format_statistics(self.h, np.asarray(v10, np.uint32), seed, dur, buf, 4096)
        return buf.value.decode()

    def decode(self, synX, synZ, p, maxit):
        ox, oz = np.zeros(self.n, np.int32), np.zeros(self.n, np.int32)
        code = self.L.qref_decode(self.dec, np.ascontiguousarray(synX, np.int32), np.ascontiguousarray(synZ, np.int32),
                                  p, maxit, ox, oz)
        return code, ox.astype(np.uint8), oz.astype(np.uint8)

    def syndrome(self, side, err):
        out = np.zeros(self.m[side], np.int32)
        self.L.qref_syndrome(self.h, side, np.ascontiguousarray(err, np.int32), out)
        return out.astype(np.uint8)

    def check_logical(self, err2n):
        return self.L.qref_check_logical(self.h, np.ascontiguousarray(err2n, np.int32))

    def bp_trace(self, side, syn, p, maxit, E, cap=None):
        cap = cap or maxit
        q = np.zeros((cap, E), np.float32)
        r = np.zeros((cap, E), np.float32)
        cv = np.full(cap, -2, np.int32)
        it = self.L.qref_bp_trace(self.dec, side, np.ascontiguousarray(syn, np.int32), p, maxit, cap, q, r, cv)
        return it, q, r, cv

    def run_frames(self, xerr, zerr, p, maxit, nthreads=0, want_out=False):
        nf = xerr.shape[0]
        k = np.zeros(8, np.uint64)
        flags = np.zeros(nf, np.uint8)
        ox = np.zeros((nf, self.n), np.uint8) if want_out else None
        oz = np.zeros((nf, self.n), np.uint8) if want_out else None
        sec = self.L.qref_run_frames(self.h, np.ascontiguousarray(xerr, np.uint8), np.ascontiguousarray(zerr, np.uint8), nf,
                                     p, maxit, nthreads, k, _opt(flags), _opt(ox), _opt(oz))
        names = ["xTested", "zTested", "corrected", "synX", "synZ", "logical", "cvX", "cvZ"]
        return dict(counters=dict(zip(names, [int(v) for v in k])), flags=flags, outX=ox, outZ=oz, seconds=sec)
