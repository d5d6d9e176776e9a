"""Tiny run of every kernel for compute-sanitizer (memcheck / racecheck): both codes, all API entry points."""
import sys
sys.path.insert(0, ".")
import numpy as np
import qec_ldpc_b200 as q

for prm, p, it, nf in [((3, 3, 6, 7, 2, 3), 0.05, 20, 300), ((4, 5, 10, 61, 9, 49), 0.06, 50, 96)]:
    code = q.Code.qc(*prm)
    dec = q.Decoder(code, 0, 64)  # several chunks
    k1 = dec.get_statistics_depolarizing(3, 0, nf, p, it)["counters"]
    x, z, sx, sz = dec.debug_generate(3, 0, nf, p)
    k2 = dec.get_stats_from_errors(x, z, p, it)["counters"]
    k3 = dec.get_stats_from_errors(x.astype(np.int32), z.astype(np.int32), p, it)["counters"]
    assert np.array_equal(k1, k2) and np.array_equal(k1, k3)
    dec.decode_batch(sx, sz, p, it)
    dec.get_statistics_weightw(5, nf, 0.02, it, 7)
    dec.debug_bp_trace(0, sx[:5], p, it, it)
    for vec in (2, 1):
        dec.configure(0, vec, 0, 0)
        dec.configure(1, vec, 0, 0)
        assert np.array_equal(dec.get_statistics_depolarizing(3, 0, nf, p, it)["counters"], k1)
    print(prm, "ok", k1.tolist())
