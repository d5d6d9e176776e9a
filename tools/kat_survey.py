#!/usr/bin/env python
"""Replays every record of the reference's published results files (tests/golden/kat_all.json) through
qldpc_get_statistics_weightw and reports which ones the CUDA path reproduces counter for counter."""
import collections
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import qec_ldpc_b200 as q  # noqa: E402
from util import CODES, COUNTERS8, golden_matrix  # noqa: E402

recs = json.load(open(os.path.join(ROOT, "tests", "golden", "kat_all.json")))
decs = {}
for code in ("C1", "C2"):
    gc = q.Code.dense(*CODES[code], golden_matrix(code, "pcmX"), golden_matrix(code, "pcmZ"), golden_matrix(code, "iMinusP"))
    decs[code] = q.Decoder(gc, 0, 1 << 17)
status = collections.Counter()
out = []
for r in recs:
    code = {42: "C1", 610: "C2"}.get(r.get("n"))
    if code is None or r.get("maxit") is None or "seed" not in r or "W" not in r:
        status["unparsed"] += 1
        out.append(dict(source=r["source"], record_index=r["record_index"], status="unparsed"))
        continue
    want = [r.get(k) for k in ["count"] + COUNTERS8]
    hit = None
    for p in dict.fromkeys([r["p_in_name"], 0.02, 0.01]):
        k = decs[code].get_statistics_weightw(r["W"], r["count"], p, r["maxit"], r["seed"])["counters"]
        got = [int(v) for v in k[:9]]
        if got == want:
            hit = p
            break
    st = "match" if hit == r["p_in_name"] else ("match_other_p" if hit is not None else "differ")
    status[(st, os.path.dirname(r["source"]))] += 1
    out.append(dict(source=r["source"], record_index=r["record_index"], status=st, p=hit, got=got, want=want))
    if st == "differ":
        print("DIFFER", r["source"][-90:], r["record_index"], "W", r["W"], "maxit", r["maxit"], "\n   got ", got, "\n   want", want, flush=True)
for k, v in sorted(status.items(), key=str):
    print(k, v)
json.dump(out, open(os.path.join(ROOT, "gpurun_out", "kat_survey.json"), "w"), indent=0)
