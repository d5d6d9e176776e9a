#!/usr/bin/env python
"""Every launch shape of the tile kernel (tile width x threads per CTA, specialised and generic instantiations) against
the CPU oracle, frame by frame, on 5000 marginal frames (p = 0.06: many frames sit close to the saturation thresholds
at their first checkpoint).  This is the check that caught the miscompiled unrolled variable loop."""
import sys
sys.path.insert(0, ".")
import numpy as np
import qec_ldpc_b200 as q
from oracle.pyoracle import Oracle

gc = q.Code.qc(4, 5, 10, 61, 9, 49)
oc = Oracle().code_qc(4, 5, 10, 61, 9, 49)
oc.set_logical(gc.dense_matrix(2))
want = oc.run_depolarizing(8, 0, 5000, 0.06, 50)
wi = want["iters"].astype(np.uint32)
dec = q.Decoder(gc, 0, 1 << 14)
bad = 0
for rep in range(2):
    for vec, threads in [(2, 128), (4, 128), (1, 128), (2, 96), (4, 64), (2, 0), (4, 0), (1, 0)]:
        for side in (0, 1):
            dec.configure(side, vec, threads, 0)
        a = dec.get_statistics_depolarizing(8, 0, 5000, 0.06, 50, per_frame=True)
        d = np.nonzero((a["iters"] != wi).any(axis=1))[0]
        ok = np.array_equal(a["counters"], want["counters"]) and np.array_equal(a["flags"], want["flags"]) and len(d) == 0
        bad += not ok
        print(rep, "vec", vec, "threads", threads, "identical to the oracle" if ok else "DIFFERS: frames %s" % d[:8], flush=True)
sys.exit(1 if bad else 0)
