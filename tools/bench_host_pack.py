#!/usr/bin/env python
"""e2e throughput of qldpc_get_stats_from_errors_i32 / _u8 (host patterns in, counters out) against the number of
host packing threads (0 = raw rows over the link, packed on the device).  One JSON line per setting."""
import json
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import qec_ldpc_b200 as q  # noqa: E402


def main():
    import torch
    frames = int(sys.argv[1]) if len(sys.argv) > 1 else 1_000_000
    code = q.Code.qc(4, 5, 10, 61, 9, 49)
    dec = q.Decoder(code, 0, frames)
    n = code.n
    x, z, _, _ = dec.debug_generate(7, 0, frames, 0.05)
    ref = None
    for dtype, name in ((np.int32, "i32"), (np.uint8, "u8")):
        for pinned in (True, False):
            hx = torch.empty((frames, n), dtype=torch.int32 if dtype == np.int32 else torch.uint8, pin_memory=pinned)
            hz = torch.empty_like(hx, pin_memory=pinned)
            hx.numpy()[:] = x
            hz.numpy()[:] = z
            for threads in (0, 2, 4, 6, 8, 10, 12, 14, 16):
                dec.set_host_threads(threads)
                fn = dec.get_stats_from_errors_ptr
                k = fn(hx.data_ptr(), hz.data_ptr(), frames, 0.05, 50, dtype().itemsize)
                torch.cuda.synchronize()
                t0 = time.perf_counter()
                reps = 2
                for _ in range(reps):
                    k = fn(hx.data_ptr(), hz.data_ptr(), frames, 0.05, 50, dtype().itemsize)
                torch.cuda.synchronize()
                dt = (time.perf_counter() - t0) / reps
                if ref is None:
                    ref = k
                assert np.array_equal(np.asarray(k), np.asarray(ref)), "counters differ"
                print(json.dumps({"layout": name, "pinned": pinned, "host_threads": threads, "frames": frames,
                                  "ms": dt * 1e3, "frames_per_s": frames / dt}), flush=True)


if __name__ == "__main__":
    main()
