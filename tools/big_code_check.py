"""A code beyond the shared-memory tier (J4K4L8, P=2053: 65 696 edges per side): the decoder takes the HBM-resident
path on its own; compared frame by frame with the CPU oracle, then timed on a 4096-frame batch."""
import sys, time
import numpy as np
sys.path.insert(0, ".")
sys.path.insert(0, "tests")
import qec_ldpc_b200 as q
from oracle.pyoracle import Oracle
O = Oracle()
prm = (4, 4, 8, 2053, 244, 2)
t0 = time.time(); gc = q.Code.qc(*prm); print("code", time.time() - t0)
dec = q.Decoder(gc, 0, 4096)
print(dec.launch_info(0), dec.launch_info(1))
oc = O.code_qc(*prm); oc.set_logical(gc.dense_matrix(2))
t0 = time.time(); b = oc.run_depolarizing(5, 0, 40, 0.03, 50); print("oracle", time.time() - t0)
a = dec.get_statistics_depolarizing(5, 0, 40, 0.03, 50, per_frame=True)
print(np.array_equal(a["counters"], b["counters"]), np.array_equal(a["flags"], b["flags"]), np.array_equal(a["iters"], b["iters"].astype(np.uint32)), a["counters"])
import torch
for nf in (4096,):
    dec.get_statistics_depolarizing(5, 0, nf, 0.03, 50)
    torch.cuda.synchronize(); t0 = time.time()
    k = dec.get_statistics_depolarizing(5, 0, nf, 0.03, 50)["counters"]
    torch.cuda.synchronize(); dt = time.time() - t0
    eu = int(k[9]) * gc.EX + int(k[10]) * gc.EZ
    print("P=2053 global: %d frames %.3fs  %.3g edge-updates/s (%.2f of HBM roofline)" % (nf, dt, eu / dt, eu * 16 / dt / 6549e9))
