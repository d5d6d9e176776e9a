"""The C-ABI library loads, exports every symbol include/qldpc_b200.h declares, and has no CPU decode path."""
import ctypes
import os
import re

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def declared_symbols():
    text = open(os.path.join(ROOT, "include", "qldpc_b200.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(qldpc_[a-z0-9_]+)\s*\(", text)))


def test_header_symbols_exported(qldpc):
    lib = ctypes.CDLL(qldpc.LIB_PATH)
    names = declared_symbols()
    assert len(names) >= 25
    for n in names:
        assert hasattr(lib, n), "missing export " + n


def test_header_cites_reference_interfaces():
    text = open(os.path.join(ROOT, "include", "qldpc_b200.h")).read()
    for cite in ["Decoder.h:40", "DecoderCPU.h:392", "DecoderGPU.h:193", "Quantum_LDPC_Code.h:26", "QEC_LDPC_CSS.cu:37"]:
        assert cite in text


def test_no_gpu_fails_loudly(qldpc):
    """Without a CUDA device the decoder refuses to exist: no silent CPU fallback."""
    import torch
    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    code = qldpc.Code.qc(3, 3, 6, 7, 2, 3)
    with pytest.raises(qldpc.QldpcError) as e:
        qldpc.Decoder(code)
    assert e.value.code == qldpc.ERR_NO_DEVICE


def test_product_does_not_touch_oracle():
    """The product package must not import, link or execute anything under oracle/."""
    pkg = os.path.join(ROOT, "qec_ldpc_b200")
    for base, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".cpp", ".h")):
                src = open(os.path.join(base, f), errors="ignore").read()
                assert "pyoracle" not in src and "liboracle" not in src and "oracle/" not in src.replace("oracle/oracle.c restates", ""), f


def test_host_packer_matches_numpy(qldpc):
    """The host-side marshalling (csrc/host_pack.cpp) against numpy, for ragged widths and thread counts; no GPU."""
    rng = np.random.default_rng(3)
    for cols in (1, 31, 32, 33, 42, 610):
        for dtype in (np.uint8, np.int32):
            a = (rng.random((131, cols)) < 0.3).astype(dtype)
            if dtype == np.int32:
                a *= rng.integers(-7, 1 << 20, size=a.shape, dtype=np.int32) | 1  # any non-zero value is a set bit
            nz = (a != 0)
            ref = np.zeros((131, (cols + 31) // 32), np.uint32)
            for c in range(cols):
                ref[:, c >> 5] |= nz[:, c].astype(np.uint32) << np.uint32(c & 31)
            for threads in (1, 2, 5):
                w = qldpc.host_pack(a, threads)
                assert np.array_equal(w, ref)
                assert np.array_equal(qldpc.host_unpack(w, cols, threads), nz.astype(np.uint8))


@pytest.mark.parametrize("n,weight,seed", [(610, 30, 2719323735), (42, 5, 1), (3, 7, 99), (4072, 45, 123456789),
                                           (6000000, 20000, 7)])
def test_weightw_stream_matches_oracle(qldpc, oracle, n, weight, seed):
    """The library's weight-W error stream (own MT19937 + MSVC uniform_int mapping, serial producer / parallel
    consumers) against the oracle's restatement of DecoderCPU.h:394-396,446-459, frame by frame; no GPU.  n = 3 makes
    the index range hit the accept-everything case of the mapping, n = 610 / 42 the rejection case; with n = 6e6 one
    index draw in 860 is rejected (~185 rejections in this run), which exercises the word-by-word path of the producer."""
    nf = 30011 if n < 100000 else 8  # several 8192-frame blocks through the producer / mapper pipeline
    ox, oz = oracle.weightw_stream(seed, weight, n, nf)
    for threads in (1, 4):
        x, z = qldpc.weightw_patterns(seed, weight, n, nf, threads)
        assert np.array_equal(qldpc.host_unpack(x, n), ox) and np.array_equal(qldpc.host_unpack(z, n), oz)
