#!/usr/bin/env python
"""Secondary measurements on the other BASELINE.json configurations (the driver's bench line is bench.py, config 2):
  C1  J3K3L6P7   p=0.05  20 iterations   (the reference's CPU-sized case)
  C4  J4K5L10P61 p=0.01 200 iterations   (early-exit divergence: almost all frames stop at 11, stragglers run to 200)
  C5  J4K4L8P509 p=0.03  50 iterations   (large code: one frame-side = 65 KB of messages, still shared-memory resident)
A name with a trailing "g" (C5g, C2g) forces the HBM-resident path (bp_global.cu) on the same configuration.
One JSON line per configuration: frames/s, edge-updates/s, mean iterations, roofline fractions, launch shapes."""
import json
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import qec_ldpc_b200 as q  # noqa: E402

CONFIGS = {"C1": ((3, 3, 6, 7, 2, 3), 0.05, 20, 4_000_000), "C2": ((4, 5, 10, 61, 9, 49), 0.05, 50, 1_000_000),
           "C4": ((4, 5, 10, 61, 9, 49), 0.01, 200, 1_000_000), "C5": ((4, 4, 8, 509, 208, 2), 0.03, 50, 100_000)}


def main():
    import torch
    names = sys.argv[1:] or ["C1", "C4", "C5"]
    peaks = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "MEASURED_PEAKS.json")
    hbm = float(json.load(open(peaks))["hbm_gbs"]) if os.path.exists(peaks) else 6650.0
    for name in names:
        forced_global = name.endswith("g")
        prm, p, maxit, frames = CONFIGS[name.rstrip("g")]
        code = q.Code.qc(*prm)
        dec = q.Decoder(code, 0, frames)
        if forced_global:
            for side in (0, 1):
                dec.configure(side, -1, 0, 0)
        stream = torch.cuda.Stream()
        dec.set_stream(stream.cuda_stream)
        for s in range(3):
            dec.get_statistics_depolarizing(7, (100 + s) * frames, frames, p, maxit)
        dec.get_timing(reset=True)
        dec.enable_timing(True)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        k = np.zeros(q.NUM_COUNTERS, np.uint64)
        steps = 3
        torch.cuda.synchronize()
        e0.record(stream)
        for s in range(steps):
            k += dec.get_statistics_depolarizing(7, s * frames, frames, p, maxit)["counters"]
        e1.record(stream)
        torch.cuda.synchronize()
        sec = e0.elapsed_time(e1) * 1e-3
        ms, _ = dec.get_timing()
        eu = int(k[9]) * code.EX + int(k[10]) * code.EZ
        bp = (ms["bp_x"] + ms["bp_z"]) * 1e-3
        ach = eu * 16 / bp / 1e9
        print(json.dumps({"config": name, "code": code.name(), "p": p, "max_iterations": maxit, "frames_per_step": frames,
                          "frames_per_s": int(k[0]) / sec, "edge_updates_per_s": eu / sec,
                          "bp_kernel_edge_updates_per_s": eu / bp, "mean_iterations": [int(k[9]) / int(k[0]), int(k[10]) / int(k[0])],
                          "frame_error_rate": 1 - int(k[3]) / int(k[0]),
                          "roofline_hbm_frac": ach / hbm, "roofline_smem_frac": ach / (148 * 128 * 1.965),
                          "kernel_ms_per_step": {a: b / steps for a, b in ms.items()},
                          "launch": [dec.launch_info(0), dec.launch_info(1)]}), flush=True)


if __name__ == "__main__":
    main()
