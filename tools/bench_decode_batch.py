#!/usr/bin/env python
"""Throughput of the host-supplied-syndrome entry point (qldpc_decode_batch, the batched Decoder::Decode): host syndromes
in (one byte per check), decisions + ErrorCode out (one byte per qubit), H2D / D2H inside the timed region."""
import json
import sys
import time

import numpy as np

sys.path.insert(0, ".")
import qec_ldpc_b200 as q  # noqa: E402

n_frames = int(sys.argv[1]) if len(sys.argv) > 1 else 1_000_000
code = q.Code.qc(4, 5, 10, 61, 9, 49)
dec = q.Decoder(code, 0, 1 << 18)
sx = np.zeros((n_frames, code.mX), np.uint8)
sz = np.zeros((n_frames, code.mZ), np.uint8)
for off in range(0, n_frames, 100_000):
    cnt = min(100_000, n_frames - off)
    _, _, a, b = dec.debug_generate(3, off, cnt, 0.05)
    sx[off:off + cnt], sz[off:off + cnt] = a, b
dec.decode_batch(sx[:1 << 18], sz[:1 << 18], 0.05, 50)
t0 = time.perf_counter()
ox, oz, fl, it = dec.decode_batch(sx, sz, 0.05, 50)
sec = time.perf_counter() - t0
print(json.dumps({"api": "qldpc_decode_batch (pageable host buffers)", "frames": n_frames, "seconds": sec,
                  "frames_per_s": n_frames / sec, "h2d_bytes": int(sx.nbytes + sz.nbytes),
                  "d2h_bytes": int(ox.nbytes + oz.nbytes + fl.nbytes + it.nbytes),
                  "syndrome_failures": int(((fl & 3) != 0).sum())}))

import torch  # noqa: E402  (pinned host memory only)
psx, psz = torch.from_numpy(sx).pin_memory(), torch.from_numpy(sz).pin_memory()
pox = torch.empty((n_frames, code.n), dtype=torch.uint8, pin_memory=True)
poz = torch.empty((n_frames, code.n), dtype=torch.uint8, pin_memory=True)
pfl = torch.empty(n_frames, dtype=torch.uint8, pin_memory=True)
pit = torch.empty((n_frames, 2), dtype=torch.int32, pin_memory=True)
args = (psx.data_ptr(), psz.data_ptr(), n_frames, 0.05, 50, pox.data_ptr(), poz.data_ptr(), pfl.data_ptr(), pit.data_ptr())
dec.decode_batch_ptr(*args)
t0 = time.perf_counter()
for _ in range(3):
    dec.decode_batch_ptr(*args)
sec = (time.perf_counter() - t0) / 3
assert np.array_equal(pox.numpy(), ox) and np.array_equal(poz.numpy(), oz) and np.array_equal(pfl.numpy(), fl)
assert np.array_equal(pit.numpy().astype(np.uint32), it)
print(json.dumps({"api": "qldpc_decode_batch (pinned host buffers)", "frames": n_frames, "seconds": sec,
                  "frames_per_s": n_frames / sec}))
