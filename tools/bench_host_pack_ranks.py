#!/usr/bin/env python
"""Aggregate host packing rate with R processes x T threads packing at once (what R ranks on one box do), for the
AVX-512 and AVX2 packers: python tools/bench_host_pack_ranks.py R T.  No GPU needed."""
import json
import multiprocessing as mp
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))


def worker(rank, threads, isa, barrier, q):
    os.environ["QLDPC_HOST_ISA"] = isa
    import qec_ldpc_b200 as ql
    rows, n = 300_000, 610
    x = (np.random.default_rng(rank).random((rows, n)) < 0.05).astype(np.int32)
    ql.host_pack(x[:1000], threads)
    best = 1e9
    for _ in range(3):
        barrier.wait()
        t0 = time.perf_counter()
        ql.host_pack(x, threads)
        best = min(best, time.perf_counter() - t0)
    barrier.wait()
    rd = ql.host_read_gbs(x.ctypes.data, x.nbytes, threads, 2)
    q.put((x.nbytes / best / 1e9, rd))


if __name__ == "__main__":
    R, T = int(sys.argv[1]), int(sys.argv[2])
    ctx = mp.get_context("spawn")
    for isa in ("avx512", "avx2"):
        barrier, q = ctx.Barrier(R), ctx.Queue()
        ps = [ctx.Process(target=worker, args=(r, T, isa, barrier, q)) for r in range(R)]
        [p.start() for p in ps]
        res = [q.get() for _ in ps]
        [p.join() for p in ps]
        print(json.dumps({"ranks": R, "threads_per_rank": T, "isa": isa, "pack_int32_gbs_sum": sum(r[0] for r in res),
                          "read_gbs_sum": sum(r[1] for r in res)}), flush=True)
