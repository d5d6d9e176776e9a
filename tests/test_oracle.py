"""CPU tests: the plain-C restatement (oracle/oracle.c) against the golden fixtures made from the unmodified
reference (tests/golden/make_golden.py), against the reference's published results files (K1..K5), and -- when
oracle/_ref is present -- live against the unmodified reference on fresh random inputs."""
import hashlib

import numpy as np
import pytest

from util import (CODES, COUNTERS8, REF_FILES, canon_bits, golden, golden_matrix, golden_params, oracle_code,
                  same_floats, unpack_rows)


def test_philox_known_answers(oracle):
    # Random123 kat_vectors, philox4x32 10 rounds
    kat = [((0, 0, 0, 0), (0, 0), (0x6627e8d5, 0xe169c58d, 0xbc57ac4c, 0x9b00dbd8)),
           ((0xffffffff,) * 4, (0xffffffff,) * 2, (0x408f276d, 0x41c83b0e, 0xa20bc7c6, 0x6d5451fd)),
           ((0x243f6a88, 0x85a308d3, 0x13198a2e, 0x03707344), (0xa4093822, 0x299f31d0),
            (0xd16cfe09, 0x94fdcceb, 0x5001e420, 0x24126ea1))]
    for ctr, key, want in kat:
        assert tuple(int(v) for v in oracle.philox(ctr, key)) == want


def test_depolarizing_thresholds(oracle):
    assert oracle.thresholds(0.0).tolist() == [0, 0, 0]
    t = oracle.thresholds(0.05)
    T = int(np.float64(np.float32(0.05)) * 4294967296.0)
    assert t.tolist() == [T // 3, 2 * T // 3, T]
    assert oracle.thresholds(1.0).tolist()[2] == 0xFFFFFFFF


def test_depolarizing_rate(oracle):
    oc = oracle_code(oracle, "C2", None)
    x = np.zeros(610, np.int64)
    z = np.zeros(610, np.int64)
    both = 0
    nf = 4000
    for f in range(nf):
        a, b = oc.depolarizing(5, f, 0.05)
        x += a
        z += b
        both += int((a & b).sum())
    tot = nf * 610
    # P(x bit) = P(z bit) = 2p/3, P(both) = p/3
    assert abs(x.sum() / tot - 2 * 0.05 / 3) < 4 * np.sqrt(0.0333 / tot)
    assert abs(z.sum() / tot - 2 * 0.05 / 3) < 4 * np.sqrt(0.0333 / tot)
    assert abs(both / tot - 0.05 / 3) < 4 * np.sqrt(0.0167 / tot)


@pytest.mark.parametrize("code", ["C1", "C2"])
def test_construction_matches_code_files(oracle, code):
    """QEC_LDPC_CSS.cu:37-131 formulas reproduce pcmX / pcmZ of the shipped code files."""
    prm = golden_params(code)
    assert prm == CODES[code]
    oc = oracle.code_qc(*prm)
    assert np.array_equal(oc.dense(0), golden_matrix(code, "pcmX"))
    assert np.array_equal(oc.dense(1), golden_matrix(code, "pcmZ"))
    # H_X H_Z^T = 0 (mod 2)
    assert not ((oc.dense(0) @ oc.dense(1).T) % 2).any()


def test_exponents_survey_values(oracle):
    hc, hd = oracle.exponents(*CODES["C2"])
    assert hc[0].tolist() == [1, 9, 20, 58, 34, 42, 12, 47, 57, 25]
    assert hd[0].tolist() == [19, 49, 14, 4, 36, 60, 52, 41, 3, 27]


def test_weightw_stream_is_mt19937(oracle):
    # std::mt19937 known answer: the 10000th output for the default seed 5489 is 4123659995
    # (ISO C++ [rand.predef]); exercised through a modulus-1 index draw that never rejects.
    x, z = oracle.weightw_stream(1, 3, 42, 5)
    assert x.shape == (5, 42) and ((x | z).sum(axis=1) <= 3).all() and ((x | z).sum(axis=1) >= 1).all()


@pytest.mark.parametrize("kat", ["K1", "K1b", "K2", "K4a", "K4b", "K5"])
def test_published_results_files(oracle, kat):
    """Known-answer tests: the reference's checked-in results (seed + counters) are reproduced bit-exactly."""
    r = golden("kat_results.json")[kat]
    oc = oracle_code(oracle, r["code"])
    k = oc.get_statistics_weightw(r["W"], r["count"], r["errorProbability"], r["maxit"], r["seed"], 0)
    got = dict(zip(["count"] + COUNTERS8, [int(v) for v in k[:9]]))
    assert got == {key: r[key] for key in got}


def test_published_results_files_short_records(oracle):
    """The oracle against every third 1000-frame, 100-iteration record the reference published for the n=610 code
    (W = 1, 4, 7, ... 58; tests/golden/kat_all.json) plus the two lightest n=42 records.  The GPU test
    test_all_published_results_files replays all 188 records; this keeps the CPU suite within minutes."""
    recs = [r for r in golden("kat_all.json") if r["group"] == "current"]
    short = sorted((r for r in recs if r["count"] == 1000 and r["maxit"] == 100), key=lambda r: r["W"])
    pick = short[::3] + [r for r in recs if r["n"] == 42 and r["W"] <= 2 and r["maxit"] == 100]
    assert len(pick) >= 20
    for r in pick:
        oc = oracle_code(oracle, {42: "C1", 610: "C2"}[r["n"]])
        k = oc.get_statistics_weightw(r["W"], r["count"], r["errorProbability"], r["maxit"], r["seed"], 0)
        got = [int(v) for v in k[:9]]
        assert got == [r[key] for key in ["count"] + COUNTERS8], (r["source"], r["record_index"])


@pytest.mark.slow
def test_published_results_file_K3(oracle):
    r = golden("kat_results.json")["K3"]
    oc = oracle_code(oracle, r["code"])
    k = oc.get_statistics_weightw(r["W"], r["count"], r["errorProbability"], r["maxit"], r["seed"], 0)
    got = dict(zip(["count"] + COUNTERS8, [int(v) for v in k[:9]]))
    assert got == {key: r[key] for key in got}


@pytest.mark.parametrize("case", ["C1a", "C2a", "C2b"])
def test_frames_against_reference_outputs(oracle, case):
    g = golden("ref_depolarizing.npz")
    code = case[:2]
    oc = oracle_code(oracle, code)
    seed, nf, maxit = [int(v) for v in g[case + "_meta"]]
    p = float(g[case + "_p"][0])
    xs, zs = unpack_rows(g[case + "_xerr"], oc.n), unpack_rows(g[case + "_zerr"], oc.n)
    for f in range(nf):  # the generator itself
        x, z = oc.depolarizing(seed, f, p)
        assert np.array_equal(x, xs[f]) and np.array_equal(z, zs[f])
    out = oc.run_frames(xs, zs, p, maxit, want_out=True)
    assert np.array_equal(out["flags"] & 63, g[case + "_flags"])
    assert np.array_equal(out["outX"], unpack_rows(g[case + "_outX"], oc.n))
    assert np.array_equal(out["outZ"], unpack_rows(g[case + "_outZ"], oc.n))
    assert [int(v) for v in out["counters"][1:9]] == g[case + "_counters"].tolist()


@pytest.mark.parametrize("code", ["C1", "C2"])
def test_traces_against_reference(oracle, code):
    g = golden("ref_traces.npz")
    oc = oracle_code(oracle, code, None)
    meta = [int(v) for v in g[code + "_meta"]]
    seed, maxit, frames = meta[0], meta[1], meta[2:]
    p = float(g[code + "_p"][0])
    for f in frames:
        x, z = oc.depolarizing(seed, f, p)
        for side, e in ((0, x), (1, z)):
            key = "%s_f%d_s%d" % (code, f, side)
            it, q, r, qt, rt = oc.bp(side, oc.syndrome(side, e), p, maxit, trace=maxit)
            assert it == int(g[key + "_iters"][0])
            keep = g[key + "_keep"]
            assert same_floats(qt[keep], g[key + "_q"]) and same_floats(rt[keep], g[key + "_r"])
            assert hashlib.sha256(canon_bits(qt[:it]).tobytes()).digest() == g[key + "_sha_q"].tobytes()
            assert hashlib.sha256(canon_bits(rt[:it]).tobytes()).digest() == g[key + "_sha_r"].tobytes()


def test_iminusp_kernel_is_rowspace(oracle):
    """iMinusP is block diagonal and its kernel is rowspace(pcmX) (+) rowspace(pcmZ)."""
    from util import gf2_rank
    imp = golden_matrix("C1", "iMinusP")
    n = 42
    assert not imp[:n, n:].any() and not imp[n:, :n].any()
    for blk, key in ((imp[:n, :n], "pcmX"), (imp[n:, n:], "pcmZ")):
        H = golden_matrix("C1", key)
        assert not ((blk @ H.T) % 2).any()                    # every row of H is in the kernel
        assert gf2_rank(blk) == n - gf2_rank(H)               # and the kernel is no larger


# ---- live comparison with the unmodified reference (when oracle/_ref is present) ------------------------

@pytest.mark.parametrize("code,p,maxit,nf", [("C1", 0.08, 20, 300), ("C2", 0.06, 50, 60), ("C2", 0.01, 31, 40)])
def test_live_reference_frames(oracle, reference, code, p, maxit, nf):
    import os
    if not os.path.exists(REF_FILES[code]):
        pytest.skip("reference code files not present on this machine")
    rc = reference.code_from_file(REF_FILES[code])
    oc = oracle.code_qc(*CODES[code])
    oc.set_logical(rc.dense(2))
    xs = np.zeros((nf, oc.n), np.uint8)
    zs = np.zeros((nf, oc.n), np.uint8)
    for f in range(nf):
        xs[f], zs[f] = oc.depolarizing(31337, 1000 + f, p)
    a = oc.run_frames(xs, zs, p, maxit, want_out=True)
    b = rc.run_frames(xs, zs, p, maxit, want_out=True)
    assert np.array_equal(a["flags"] & 63, b["flags"])
    assert np.array_equal(a["outX"], b["outX"]) and np.array_equal(a["outZ"], b["outZ"])
    for f in range(3):
        for side, e in ((0, xs[f]), (1, zs[f])):
            syn = oc.syndrome(side, e)
            assert np.array_equal(syn, rc.syndrome(side, e))
            it_r, q_r, r_r, _ = rc.bp_trace(side, syn, p, maxit, oc.E[side])
            it_o, _, _, qt, rt = oc.bp(side, syn, p, maxit, trace=maxit)
            assert it_r == it_o and same_floats(qt[:it_o], q_r[:it_r]) and same_floats(rt[:it_o], r_r[:it_r])
