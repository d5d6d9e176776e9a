#!/usr/bin/env python
"""Host-side marshalling rates on this box: the packer's throughput (one int32 / one byte per bit -> packed words)
against its thread count, beside the streaming-read bandwidth the same threads reach (the host-memory roofline of the
int32 layout).  No GPU needed.  One JSON line per thread count."""
import json
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import qec_ldpc_b200 as q  # noqa: E402

rows, n = 400_000, 610
x32 = (np.random.default_rng(1).random((rows, n)) < 0.05).astype(np.int32)
x8 = x32.astype(np.uint8)
for threads in (1, 2, 4, 8, 16):
    if threads > (os.cpu_count() or 1):
        break
    out = {"threads": threads, "rows": rows, "cols": n}
    for name, a in (("int32", x32), ("uint8", x8)):
        q.host_pack(a[:1000], threads)
        best = 1e9
        for _ in range(3):
            t0 = time.perf_counter()
            q.host_pack(a, threads)
            best = min(best, time.perf_counter() - t0)
        out["pack_%s_gbs" % name] = a.nbytes / best / 1e9
        out["pack_%s_mframes_s" % name] = rows / best / 1e6
    out["read_gbs"] = q.host_read_gbs(x32.ctypes.data, x32.nbytes, threads, 3)
    print(json.dumps(out), flush=True)
