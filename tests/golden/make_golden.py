"""Generates the golden fixtures in this directory from the UNMODIFIED reference.

Run in the authoring container (needs /root/reference and oracle/_ref/libqldpc_ref.so, built by `make -C oracle ref`):

    python tests/golden/make_golden.py

Outputs (committed):
  kat_results.json       K1..K5: seed/W/COUNT/MAXIT/p and the published counters, parsed from the reference's results
                         files (QEC_LDPC/results/...), see SURVEY.md section 4.
  codes.npz              both shipped code files, bit-packed: pcmX, pcmZ, iMinusP (lines 2-4) + parameters (line 1).
  ref_depolarizing.npz   outputs of the reference's Decode / GetSyndrome / CheckLogicalError (through
                         oracle/ref_harness.cpp:qref_run_frames) on depolarizing patterns of our Philox generator.
  ref_traces.npz         per-iteration messages of the reference's EqNodeUpdate / VarNodeUpdate
                         (oracle/ref_harness.cpp:qref_bp_trace).
The input patterns are produced by oracle.c's Philox generator (the generator is ours, not the reference's); every
OUTPUT stored here comes from reference code.
"""
import hashlib
import json
import os
import re
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
from oracle.pyoracle import Oracle, Reference, build  # noqa: E402

REF = "/root/reference"
FILES = {"C1": REF + "/J_3_K_3_L_6_P_7_s_2_t_3.txt", "C2": REF + "/QEC_LDPC/code610.txt"}
RES = REF + "/QEC_LDPC/results/"
KATS = {
    "K1": ("C1", RES + "[2,3,6,7,2,3]/[J=3,K=3,L=6,P=7,s=2,t=3][[n=42,k=0]]_W_3_MAX_1000_p_0.02.txt", 0.02, 1),
    "K1b": ("C1", RES + "[2,3,6,7,2,3]/[J=3,K=3,L=6,P=7,s=2,t=3][[n=42,k=0]]_W_3_MAX_1000_p_0.02.txt", 0.02, 0),
    "K2": ("C1", RES + "[2,3,6,7,2,3]/[J=3,K=3,L=6,P=7,s=2,t=3][[n=42,k=0]]_W_10_MAX_100_p_0.02.txt", 0.02),
    "K3": ("C2", RES + "[J=4,K=5,L=10,P=61,s=9,t=49][[n=610,k=61]]_W_15_MAX_100_p_0.01.txt", 0.01),
    "K4a": ("C2", RES + "[4,5,10,61,9,49]/[J=4,K=5,L=10,P=61,s=9,t=49][[n=610,k=61]]_W_30_MAX_100_p_0.02.txt", 0.02),
    "K4b": ("C2", RES + "[4,5,10,61,9,49]/[J=4,K=5,L=10,P=61,s=9,t=49][[n=610,k=61]]_W_45_MAX_100_p_0.02.txt", 0.02),
    # labelled p_0.01 but produced with errorProbability 0.02
    "K5": ("C2", RES + "[4,5,10,61,9,49]/[J=4,K=5,L=10,P=61,s=9,t=49][[n=610,k=61]]_W_30_MAX_100_p_0.01.txt", 0.02),
}
FIELDS = {"Rand Seed": "seed", "Errors Tested": "count", "Errors With X": "xTested", "Errors With Z": "zTested",
          "Error Weight": "W", "Corrected": "corrected", "Syndrome Errors X": "synX", "Syndrome Errors Z": "synZ",
          "Logical Errors": "logical", "Convergence Fail X": "cvX", "Convergence Fail Z": "cvZ",
          "Duration(micro-s)": "duration_us"}


def parse_record(path, index):
    """The index-th record of a results file (records are appended, CodeStatistics.h:22-37 / main.cu:100-103)."""
    recs = []
    for line in open(path).read().splitlines():
        if ":" not in line:
            continue
        k, v = line.split(":", 1)
        if k.strip() == "Code":
            recs.append({"code_name": v.strip()})
        elif k.strip() in FIELDS and recs:
            recs[-1][FIELDS[k.strip()]] = int(v)
    rec = recs[index]
    rec["maxit"] = int(re.search(r"_MAX_(\d+)_", path).group(1))
    rec["record_index"] = index
    return rec


def canon(a):
    a = np.array(a, np.float32, copy=True)
    a[np.isnan(a)] = np.float32(np.nan)
    v = a.view(np.uint32).copy()
    v[np.isnan(a)] = 0x7FC00000
    return v


def main():
    build(ref=True)
    O, R = Oracle(), Reference()
    kat = {}
    for name, spec in KATS.items():
        code, path, p = spec[:3]
        rec = parse_record(path, spec[3] if len(spec) > 3 else 0)
        rec["code"] = code
        rec["errorProbability"] = p
        rec["source"] = path[len(REF) + 1:]
        if rec["count"] <= 10000:  # keep the record's text for the results-writer tests (CodeStatistics.h:22-37)
            blocks = [b for b in open(path).read().replace("\r", "").split("Code: ") if b.strip()]
            rec["text"] = "Code: " + blocks[rec["record_index"]].rstrip("\n") + "\n"
        kat[name] = rec
    json.dump(kat, open(os.path.join(HERE, "kat_results.json"), "w"), indent=1, sort_keys=True)

    codes, dep, tr = {}, {}, {}
    cases = {"C1": [("a", 2024, 0.05, 20, 256)], "C2": [("a", 2024, 0.05, 50, 192), ("b", 77, 0.01, 200, 64)]}
    for cname, path in FILES.items():
        rc = R.code_from_file(path)
        prm = (rc.J, rc.K, rc.Lc, rc.P, rc.sigma, rc.tau)
        codes[cname + "_params"] = np.array(prm, np.int32)
        for which, key in enumerate(["pcmX", "pcmZ", "iMinusP"]):
            m = rc.dense(which)
            codes["%s_%s" % (cname, key)] = np.packbits(m.astype(np.uint8), axis=1)
            codes["%s_%s_shape" % (cname, key)] = np.array(m.shape, np.int32)
        codes[cname + "_name"] = np.frombuffer(rc.name().encode(), np.uint8)
        oc = O.code_qc(*prm)
        n = oc.n
        for tag, seed, p, maxit, nf in cases[cname]:
            xs = np.zeros((nf, n), np.uint8)
            zs = np.zeros((nf, n), np.uint8)
            for f in range(nf):
                xs[f], zs[f] = oc.depolarizing(seed, f, p)
            out = rc.run_frames(xs, zs, p, maxit, want_out=True)
            key = "%s%s" % (cname, tag)
            dep[key + "_meta"] = np.array([seed, nf, maxit], np.int64)
            dep[key + "_p"] = np.array([p], np.float32)
            dep[key + "_xerr"] = np.packbits(xs, axis=1)
            dep[key + "_zerr"] = np.packbits(zs, axis=1)
            dep[key + "_outX"] = np.packbits(out["outX"], axis=1)
            dep[key + "_outZ"] = np.packbits(out["outZ"], axis=1)
            dep[key + "_flags"] = out["flags"]
            dep[key + "_counters"] = np.array([out["counters"][k] for k in
                                               ["xTested", "zTested", "corrected", "synX", "synZ", "logical", "cvX",
                                                "cvZ"]], np.int64)
        # traces: reference EqNodeUpdate/VarNodeUpdate per iteration
        seed, p, maxit = 2024, 0.05, (20 if cname == "C1" else 50)
        frames = [0, 1, 2, 3] if cname == "C1" else [0, 5]
        for f in frames:
            x, z = oc.depolarizing(seed, f, p)
            for side, e in ((0, x), (1, z)):
                syn = rc.syndrome(side, e)
                it, q, r, cv = rc.bp_trace(side, syn, p, maxit, oc.E[side])
                key = "%s_f%d_s%d" % (cname, f, side)
                tr[key + "_iters"] = np.array([it], np.int32)
                tr[key + "_conv"] = cv[:it]
                tr[key + "_sha_q"] = np.frombuffer(hashlib.sha256(canon(q[:it]).tobytes()).digest(), np.uint8)
                tr[key + "_sha_r"] = np.frombuffer(hashlib.sha256(canon(r[:it]).tobytes()).digest(), np.uint8)
                keep = list(range(it)) if cname == "C1" else [i for i in (0, 1, 10, it - 1) if i < it]
                tr[key + "_keep"] = np.array(keep, np.int32)
                tr[key + "_q"] = q[keep]
                tr[key + "_r"] = r[keep]
        tr[cname + "_meta"] = np.array([seed, maxit] + frames, np.int64)
        tr[cname + "_p"] = np.array([p], np.float32)
    np.savez_compressed(os.path.join(HERE, "codes.npz"), **codes)
    np.savez_compressed(os.path.join(HERE, "ref_depolarizing.npz"), **dep)
    np.savez_compressed(os.path.join(HERE, "ref_traces.npz"), **tr)
    for f in sorted(os.listdir(HERE)):
        print("%8d  %s" % (os.path.getsize(os.path.join(HERE, f)), f))


if __name__ == "__main__":
    main()
