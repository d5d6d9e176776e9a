#!/usr/bin/env python
"""Randomised differential run: CUDA path vs CPU oracle, frame by frame, over many (code, p, max iterations, seed)
combinations, every tile width and the HBM-resident path with random slot-pool sizes (vec=-1).  Prints one line per case; exits non-zero on the first disagreement."""
import sys
import time

import numpy as np

sys.path.insert(0, ".")
import qec_ldpc_b200 as q  # noqa: E402
from oracle.pyoracle import Oracle  # noqa: E402

budget = float(sys.argv[1]) if len(sys.argv) > 1 else 120.0
O = Oracle()
rng = np.random.default_rng(20261018)
codes = {"C1": (3, 3, 6, 7, 2, 3), "C2": (4, 5, 10, 61, 9, 49), "C5": (4, 4, 8, 509, 208, 2),
         # further kernel shapes: (8,3)/(8,4), (6,2)/(6,3), (12,5)/(12,6), (10,3)/(10,4), and (14,3)/(14,4) with no tile kernel
         "S8": (3, 4, 8, 13, 5, 2), "S6": (2, 3, 6, 7, 2, 3), "S12": (5, 6, 12, 13, 4, 2), "S10": (3, 4, 10, 31, 2, 2),
         "S14": (3, 4, 14, 13, 3, 2)}
objs = {}
for name, prm in codes.items():
    gc = q.Code.qc(*prm)
    oc = O.code_qc(*prm)
    oc.set_logical(gc.dense_matrix(2))
    objs[name] = (gc, oc, q.Decoder(gc, 0, 1 << 15))
t_end = time.time() + budget
total = 0
case = 0
while time.time() < t_end:
    name = rng.choice(["C1", "C2", "C2", "C2", "C5", "S8", "S6", "S12", "S10", "S14"])
    gc, oc, dec = objs[name]
    p = float(np.float32(rng.choice([0.005, 0.01, 0.02, 0.03, 0.05, 0.07, 0.1, 0.2, 1e-6])))
    maxit = int(rng.choice([1, 3, 10, 11, 20, 37, 50, 100]))
    nf = {"C1": 20000, "C2": 4000, "C5": 300}.get(name, 6000)
    seed = int(rng.integers(0, 2**62))
    first = int(rng.integers(0, 2**40))
    vec = int(rng.choice([0, 4, 2, 1, -1]))
    slots = int(rng.choice([0, 32, 64, 256, 1024])) if vec < 0 else 0
    for side in (0, 1):
        try:
            dec.configure(side, vec, slots, 0)
        except q.QldpcError:
            dec.configure(side, 0, 0, 0)
    a = dec.get_statistics_depolarizing(seed, first, nf, p, maxit, per_frame=True)
    b = oc.run_depolarizing(seed, first, nf, p, maxit)
    ok = (np.array_equal(a["counters"], b["counters"]) and np.array_equal(a["flags"], b["flags"])
          and np.array_equal(a["iters"], b["iters"].astype(np.uint32)))
    total += nf
    case += 1
    print("%3d %s p=%.6g maxit=%3d vec=%d frames=%d fer=%.4f nan=%d %s" % (
        case, name, p, maxit, dec.launch_info(0)["vec"], nf, 1 - int(a["counters"][3]) / nf, int(a["counters"][11]),
        "ok" if ok else "MISMATCH"), flush=True)
    if not ok:
        sys.exit(1)
print("all %d cases identical (%d frames)" % (case, total))
