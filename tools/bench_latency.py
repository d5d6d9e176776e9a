#!/usr/bin/env python
"""Latency of the drop-in Decode call (Decoder::Decode, Decoder.h:40-43 / DecoderCPU.h:317-390) for small batches:
p50 / p99 of qldpc_decode_batch with 1, 32, 1024 and 4096 host frames (J4K5L10P61, p=0.05, 50 iterations), beside the
per-frame time of the CPU reference (oracle/_ref, one thread -- what a reference-style per-frame loop pays), and the
break-even batch size.  One JSON line -> profiles/r2/latency.json."""
import ctypes as C
import json
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import qec_ldpc_b200 as q  # noqa: E402


def main():
    code = q.Code.qc(4, 5, 10, 61, 9, 49)
    dec = q.Decoder(code, 0, 1 << 13)
    p, maxit = 0.05, 50
    _, _, sx, sz = dec.debug_generate(5, 0, 8192, p)
    rows = []
    for nf in (1, 2, 8, 32, 128, 1024, 2048, 4096):
        reps = 400 if nf <= 128 else 100
        ox, oz = np.zeros((nf, code.n), np.uint8), np.zeros((nf, code.n), np.uint8)
        fl, it = np.zeros(nf, np.uint8), np.zeros((nf, 2), np.uint32)
        ts = []
        for r in range(reps + 20):
            o = (r * nf) % (8192 - nf + 1)
            a, b = np.ascontiguousarray(sx[o:o + nf]), np.ascontiguousarray(sz[o:o + nf])
            t0 = time.perf_counter()
            dec.decode_batch_ptr(a.ctypes.data, b.ctypes.data, nf, p, maxit, ox.ctypes.data, oz.ctypes.data, fl.ctypes.data,
                                 it.ctypes.data)
            ts.append(time.perf_counter() - t0)
        ts = np.array(ts[20:]) * 1e6
        rows.append({"frames": nf, "p50_us": float(np.percentile(ts, 50)), "p99_us": float(np.percentile(ts, 99)),
                     "min_us": float(ts.min()), "us_per_frame_p50": float(np.percentile(ts, 50)) / nf})
    out = {"code": "J4K5L10P61", "p": p, "max_iterations": maxit, "api": "qldpc_decode_batch (host byte buffers, pageable)",
           "rows": rows}
    # raw launch + sync cost of an empty kernel on this box, for scale
    import torch
    torch.cuda.synchronize()
    x = torch.zeros(1, device="cuda")
    ts = []
    for _ in range(300):
        t0 = time.perf_counter()
        x.add_(1)
        torch.cuda.synchronize()
        ts.append(time.perf_counter() - t0)
    out["launch_plus_sync_us_p50"] = float(np.percentile(np.array(ts[50:]) * 1e6, 50))
    try:
        from oracle.pyoracle import Oracle, Reference
        from bench import golden_code_file, golden_matrix
        oc = Oracle().code_qc(4, 5, 10, 61, 9, 49)
        oc.set_logical(golden_matrix("C2", "iMinusP"))
        x, z = oc.depolarizing_bulk(5, 0, 256, p)
        if Reference.available():
            rc = Reference().code_from_file(golden_code_file("C2"))
            synx = np.stack([oc.syndrome(0, x[f]) for f in range(256)])
            synz = np.stack([oc.syndrome(1, z[f]) for f in range(256)])
            t0 = time.perf_counter()
            for f in range(256):
                rc.decode(synx[f], synz[f], p, maxit)
            cpu_us = (time.perf_counter() - t0) / 256 * 1e6
            out["cpu_reference_decode_us_per_frame"] = cpu_us
            out["cpu_kind"] = "reference (oracle/_ref DecoderCPU::Decode, one thread)"
            be = next((r["frames"] for r in rows if r["p50_us"] < cpu_us * r["frames"]), None)
            out["break_even_frames"] = be
    except Exception as e:  # the checker is optional here
        out["cpu_reference_error"] = str(e)
    print(json.dumps(out))


if __name__ == "__main__":
    main()
