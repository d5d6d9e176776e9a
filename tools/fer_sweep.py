#!/usr/bin/env python
"""FER-vs-p threshold sweep (BASELINE config 3), sharded over all ranks of a torchrun launch.

  python tools/fer_sweep.py [--code code610.txt | --qc 4,5,10,61,9,49] [--p 0.01:0.10:10] [--frames 1000000]
  python -m torch.distributed.run --nproc-per-node 8 --master-addr 127.0.0.1 tools/fer_sweep.py ...

Every p-point decodes the same global frame-id range split contiguously over the ranks (results are identical for any
rank count); one all-reduce of the counters per point.  Output: one JSON line per point with the frame error rate and
its Wilson 95% interval, the logical / syndrome-failure split and the mean executed iterations."""
import argparse
import json
import math
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))


def wilson(k, n, z=1.96):
    if n == 0:
        return 0.0, 1.0
    ph = k / n
    d = 1 + z * z / n
    c = ph + z * z / (2 * n)
    h = z * math.sqrt(ph * (1 - ph) / n + z * z / (4 * n * n))
    return max(0.0, (c - h) / d), min(1.0, (c + h) / d)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--code", default=None, help="code file in the reference's 4-line format")
    ap.add_argument("--qc", default="4,5,10,61,9,49", help="J,K,L,P,sigma,tau when no --code is given")
    ap.add_argument("--p", default="0.01:0.10:10", help="lo:hi:points (inclusive, linear) or a comma list")
    ap.add_argument("--frames", type=int, default=1_000_000, help="frames per p-point and batch (whole job)")
    ap.add_argument("--target-errors", type=int, default=0,
                    help="stopping rule: keep adding batches of --frames until this many frame errors were seen (0 = one batch)")
    ap.add_argument("--max-frames", type=int, default=100_000_000, help="cap on frames per p-point under --target-errors")
    ap.add_argument("--max-iterations", type=int, default=50)
    ap.add_argument("--seed", type=int, default=20261018)
    args = ap.parse_args()

    import torch
    import torch.distributed as dist
    import qec_ldpc_b200 as q
    from qec_ldpc_b200.sharding import allreduce_max, run_sharded

    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    rank = int(os.environ.get("RANK", "0"))
    torch.cuda.set_device(local)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    code = q.Code.from_file(args.code) if args.code else q.Code.qc(*[int(v) for v in args.qc.split(",")])
    dec = q.Decoder(code, local, min(1 << 20, max(1, args.frames // world + 1)))
    if ":" in args.p:
        lo, hi, k = args.p.split(":")
        ps = np.linspace(float(lo), float(hi), int(k))
    else:
        ps = np.array([float(v) for v in args.p.split(",")])
    for i, p in enumerate(ps):
        p = float(np.float32(p))
        t0 = time.perf_counter()
        k = np.zeros(q.NUM_COUNTERS, np.uint64)
        while True:  # batches continue the same global frame-id stream, so the result does not depend on the batching
            done = int(k[0])
            k += run_sharded(lambda first, n: dec.get_statistics_depolarizing(args.seed + i, first, n, p,
                                                                              args.max_iterations)["counters"],
                             args.frames, first_frame=done, device="cuda")
            if int(k[0]) - int(k[3]) >= args.target_errors or int(k[0]) >= args.max_frames or args.target_errors <= 0:
                break
        torch.cuda.synchronize()
        sec = allreduce_max(time.perf_counter() - t0, "cuda")
        if rank == 0:
            n = int(k[0])
            fails = n - int(k[3])
            lo95, hi95 = wilson(fails, n)
            print(json.dumps({"code": code.name(), "p": p, "frames": n, "frame_error_rate": fails / max(n, 1),
                              "fer_ci95": [lo95, hi95], "logical": int(k[6]), "syndrome_fail_x": int(k[4]),
                              "syndrome_fail_z": int(k[5]), "convergence_fail_x": int(k[7]),
                              "convergence_fail_z": int(k[8]), "mean_iterations": [int(k[9]) / max(n, 1), int(k[10]) / max(n, 1)],
                              "nan_frames": int(k[11]), "seconds": sec, "frames_per_s": n / sec, "n_gpus": world}),
                  flush=True)
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
