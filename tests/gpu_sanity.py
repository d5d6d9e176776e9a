"""Ad-hoc GPU sanity run (not collected by pytest): parity of generator, traces and statistics + a timing."""
import sys, time
import numpy as np
sys.path.insert(0, ".")
import qec_ldpc_b200 as q
from oracle.pyoracle import Oracle

O = Oracle()


def same(a, b):
    """bit-identical, NaN payloads excepted (x86 and the GPU produce different default NaNs)"""
    na, nb = np.isnan(a), np.isnan(b)
    return np.array_equal(na, nb) and np.array_equal(a.view(np.uint32)[~na], b.view(np.uint32)[~nb])

for prm, p, maxit, nf in [((3, 3, 6, 7, 2, 3), 0.05, 20, 2000), ((4, 5, 10, 61, 9, 49), 0.05, 50, 2000)]:
    code = q.Code.qc(*prm)
    oc = O.code_qc(*prm)
    oc.set_logical(code.dense_matrix(2))
    dec = q.Decoder(code, 0, 1 << 16)
    print(prm, dec.launch_info(0), dec.launch_info(1))
    x, z, sx, sz = dec.debug_generate(7, 100, 64, p)
    for f in range(64):
        ox, oz = oc.depolarizing(7, 100 + f, p)
        assert np.array_equal(ox, x[f]) and np.array_equal(oz, z[f]), "gen"
        assert np.array_equal(oc.syndrome(0, ox), sx[f]) and np.array_equal(oc.syndrome(1, oz), sz[f]), "syn"
    print("  generator + syndrome bit-exact")
    for side, syn in ((0, sx), (1, sz)):
        qt, rt, it = dec.debug_bp_trace(side, syn[:16], p, maxit, maxit)
        for f in range(16):
            oit, _, _, oq, orr = oc.bp(side, syn[f], p, maxit, trace=maxit)
            assert oit == it[f], (side, f, oit, it[f])
            assert same(orr[:oit], rt[f, :oit]), ("r", side, f)
            assert same(oq[:oit], qt[f, :oit]), ("q", side, f)
    print("  per-iteration messages bit-exact on 16 frames/side")
    g = dec.get_statistics_depolarizing(7, 0, nf, p, maxit, per_frame=True)
    o = oc.run_depolarizing(7, 0, nf, p, maxit)
    print("  gpu", g["counters"].tolist())
    print("  cpu", o["counters"].tolist())
    assert np.array_equal(g["counters"], o["counters"])
    assert np.array_equal(g["flags"], o["flags"]) and np.array_equal(g["iters"], o["iters"])
    for nfb in (100000, 1000000):
        t = time.time(); g = dec.get_statistics_depolarizing(9, 0, nfb, p, maxit); dt = time.time() - t
        k = g["counters"]
        eu = int(k[9]) * code.EX + int(k[10]) * code.EZ
        print("  %d frames %.3fs -> %.0f frames/s, %.3e edge-updates/s, mean it %.2f/%.2f" % (nfb, dt, nfb / dt, eu / dt, k[9] / nfb, k[10] / nfb), k.tolist())
print("OK")
