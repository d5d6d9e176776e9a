"""CPU tests of the product's host logic through the C ABI: code construction (closed-form circulant edge tables),
the 4-line code-file reader / writer, the logical-check matrix, and error behaviour.  No GPU needed."""
import os

import numpy as np
import pytest

from util import CODES, REF_FILES, gf2_rank, golden, golden_matrix, golden_params


@pytest.mark.parametrize("code", ["C1", "C2"])
def test_qc_construction_matches_golden(qldpc, oracle, code):
    c = qldpc.Code.qc(*golden_params(code))
    assert np.array_equal(c.dense_matrix(0), golden_matrix(code, "pcmX"))
    assert np.array_equal(c.dense_matrix(1), golden_matrix(code, "pcmZ"))
    assert c.is_qc and c.is_css()
    assert c.name() == golden("codes.npz")[code + "_name"].tobytes().decode()
    oc = oracle.code_qc(*golden_params(code))
    for s in (0, 1):
        chk_var, var_chk, var_edge = oc.tables(s)
        assert np.array_equal(c.csr(s), chk_var)
        a, b = c.csc(s)
        assert np.array_equal(a, var_chk) and np.array_equal(b, var_edge)
        assert (np.diff(c.csr(s), axis=1) > 0).all() and (np.diff(a, axis=1) > 0).all()  # ascending, DecoderCPU.h:51-63
    hc, hd = oracle.exponents(*golden_params(code))
    assert np.array_equal(c.exponents(0), hc) and np.array_equal(c.exponents(1), hd)


def test_info_sizes(qldpc):
    c = qldpc.Code.qc(*CODES["C2"])
    assert (c.n, c.mX, c.mZ, c.dcX, c.dcZ, c.dvX, c.dvZ, c.EX, c.EZ) == (610, 244, 305, 10, 10, 4, 5, 2440, 3050)
    c5 = qldpc.Code.qc(*CODES["C5"])
    assert (c5.n, c5.mX, c5.mZ, c5.EX, c5.EZ) == (4072, 2036, 2036, 16288, 16288) and c5.is_css()


@pytest.mark.parametrize("code", ["C1", "C2"])
def test_dense_constructor_and_logical_from_file_matrix(qldpc, code):
    """Quantum_LDPC_Code(J..tau, pcmX, pcmZ, imp): tables by ascending scan equal the closed form; the row-reduced
    logical check has the same kernel as the supplied iMinusP."""
    prm = golden_params(code)
    imp = golden_matrix(code, "iMinusP")
    c = qldpc.Code.dense(*prm, golden_matrix(code, "pcmX"), golden_matrix(code, "pcmZ"), imp)
    q = qldpc.Code.qc(*prm)
    assert c.is_qc and c.logical_from_file
    for s in (0, 1):
        assert np.array_equal(c.csr(s), q.csr(s))
        assert np.array_equal(c.csc(s)[1], q.csc(s)[1])
    L = c.dense_matrix(2)
    assert L.shape[0] == gf2_rank(imp)
    assert gf2_rank(np.vstack([L, imp])) == L.shape[0]  # same row space, hence same kernel


@pytest.mark.parametrize("code", ["C1", "C2"])
def test_generated_logical_check_equivalent_to_iminusp(qldpc, code):
    """Codes built from (J,K,L,P,sigma,tau) have no iMinusP; the generated check must span the same row space."""
    c = qldpc.Code.qc(*golden_params(code))
    assert not c.logical_from_file
    L = c.dense_matrix(2)
    imp = golden_matrix(code, "iMinusP")
    r = gf2_rank(imp)
    assert L.shape[0] == r and gf2_rank(L) == r and gf2_rank(np.vstack([L, imp])) == r


def test_host_helpers_match_oracle(qldpc, oracle):
    c = qldpc.Code.qc(*CODES["C1"])
    oc = oracle.code_qc(*CODES["C1"])
    oc.set_logical(c.dense_matrix(2))
    rng = np.random.default_rng(3)
    for _ in range(50):
        e = (rng.random(84) < 0.1).astype(np.int32)
        assert np.array_equal(c.syndrome(0, e[:42]), oc.syndrome(0, e[:42].astype(np.uint8)))
        assert np.array_equal(c.syndrome(1, e[42:]), oc.syndrome(1, e[42:].astype(np.uint8)))
        assert c.check_logical(e) == bool(oc.check_logical(e.astype(np.uint8)))


def test_file_round_trip(qldpc, tmp_path):
    prm = golden_params("C1")
    imp = golden_matrix("C1", "iMinusP")
    c = qldpc.Code.dense(*prm, golden_matrix("C1", "pcmX"), golden_matrix("C1", "pcmZ"), imp)
    path = str(tmp_path / "code.txt")
    c.write_file(path)
    text = open(path).read()
    assert text.count("\n") == 3 and not text.endswith("\n")  # 4 lines, no trailing newline (as shipped)
    d = qldpc.Code.from_file(path)
    assert (d.J, d.K, d.info.L, d.P, d.sigma, d.tau) == prm
    for w in (0, 1, 2):
        assert np.array_equal(c.dense_matrix(w), d.dense_matrix(w))
    if os.path.exists(REF_FILES["C1"]):  # byte-identical to the shipped file
        assert text == open(REF_FILES["C1"]).read()
    # a generated-logical code written to a file reloads with an equivalent check
    g = qldpc.Code.qc(*prm)
    g.write_file(path)
    h = qldpc.Code.from_file(path)
    assert gf2_rank(np.vstack([g.dense_matrix(2), h.dense_matrix(2)])) == g.logical_rows == h.logical_rows


def test_reads_shipped_files_when_present(qldpc):
    for code, path in REF_FILES.items():
        if not os.path.exists(path):
            pytest.skip("reference data files not present")
        c = qldpc.Code.from_file(path)
        assert np.array_equal(c.dense_matrix(0), golden_matrix(code, "pcmX"))
        assert np.array_equal(c.dense_matrix(1), golden_matrix(code, "pcmZ"))
        assert c.is_qc and c.logical_from_file


def test_file_without_iminusp_line_never_flags_logical(qldpc, tmp_path):
    """Quantum_LDPC_Code.h:28-41: missing entries parse as zeros, so CheckLogicalError is always false."""
    c = qldpc.Code.qc(*CODES["C1"])
    path = str(tmp_path / "code.txt")
    c.write_file(path)
    lines = open(path).read().split("\n")
    open(path, "w").write("\n".join(lines[:3]))
    d = qldpc.Code.from_file(path)
    assert d.logical_rows == 0 and d.logical_from_file
    assert not d.check_logical(np.ones(84, np.int32))


def test_error_behaviour(qldpc, tmp_path):
    with pytest.raises(qldpc.QldpcError) as e:
        qldpc.Code.from_file(str(tmp_path / "missing.txt"))
    assert e.value.code == qldpc.ERR_IO and "Unable to find code file" in str(e.value)  # Quantum_LDPC_Code.h:78
    with pytest.raises(qldpc.QldpcError):
        qldpc.Code.qc(4, 5, 10, 62, 2, 3)  # sigma not invertible mod P
    bad = golden_matrix("C1", "pcmX").copy()
    bad[0, np.nonzero(bad[0])[0][0]] = 0  # irregular row
    with pytest.raises(qldpc.QldpcError):
        qldpc.Code.dense(*golden_params("C1"), bad, golden_matrix("C1", "pcmZ"))


def test_non_css_parameters_detected(qldpc):
    # J4K5L10 is not constructible at P=509: H_X H_Z^T != 0
    c = qldpc.Code.qc(4, 5, 10, 509, 208, 2)
    assert not c.is_css()
