"""Shared helpers for the test-suite (golden fixture loading, comparisons)."""
import json
import os

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
GOLDEN = os.path.join(HERE, "golden")
REF_ROOT = "/root/reference"
REF_FILES = {"C1": REF_ROOT + "/J_3_K_3_L_6_P_7_s_2_t_3.txt", "C2": REF_ROOT + "/QEC_LDPC/code610.txt"}
CODES = {"C1": (3, 3, 6, 7, 2, 3), "C2": (4, 5, 10, 61, 9, 49), "C5": (4, 4, 8, 509, 208, 2)}
COUNTERS8 = ["xTested", "zTested", "corrected", "synX", "synZ", "logical", "cvX", "cvZ"]

_cache = {}


def golden(name):
    if name not in _cache:
        path = os.path.join(GOLDEN, name)
        _cache[name] = json.load(open(path)) if name.endswith(".json") else np.load(path)
    return _cache[name]


def golden_matrix(code, key):
    g = golden("codes.npz")
    shp = g["%s_%s_shape" % (code, key)]
    return np.unpackbits(g["%s_%s" % (code, key)], axis=1)[:, :shp[1]].astype(np.int32)


def golden_params(code):
    return tuple(int(v) for v in golden("codes.npz")[code + "_params"])


def unpack_rows(packed, ncols):
    return np.unpackbits(packed, axis=1)[:, :ncols]


def same_floats(a, b):
    """Bit-identical float arrays, NaN payloads excepted (x86 and the GPU generate different default NaNs)."""
    a = np.ascontiguousarray(a, np.float32)
    b = np.ascontiguousarray(b, np.float32)
    na, nb = np.isnan(a), np.isnan(b)
    return a.shape == b.shape and np.array_equal(na, nb) and np.array_equal(a.view(np.uint32)[~na], b.view(np.uint32)[~nb])


def canon_bits(a):
    a = np.ascontiguousarray(a, np.float32)
    v = a.view(np.uint32).copy()
    v[np.isnan(a)] = 0x7FC00000
    return v


def gf2_rank(m):
    m = (np.array(m, dtype=np.uint8) & 1).copy()
    rows, cols = m.shape
    r = 0
    for c in range(cols):
        piv = np.nonzero(m[r:, c])[0]
        if piv.size == 0:
            continue
        p = r + piv[0]
        if p != r:
            m[[r, p]] = m[[p, r]]
        others = np.nonzero(m[:, c])[0]
        others = others[others != r]
        m[others] ^= m[r]
        r += 1
        if r == rows:
            break
    return r


def oracle_code(oracle, code, logical="golden"):
    """Oracle code object for a named configuration with its logical-check matrix set."""
    oc = oracle.code_qc(*CODES[code])
    if isinstance(logical, str) and logical == "golden":
        oc.set_logical(golden_matrix(code, "iMinusP"))
    elif logical is not None:
        oc.set_logical(logical)
    return oc
