"""Throughput of qldpc_get_statistics_weightw (the reference's GetStatistics(W, COUNT, p, MAXIT, seed) with its serial
mt19937 error stream) on J4K5L10P61, 1M frames per weight."""
import sys, time
sys.path.insert(0, ".")
import qec_ldpc_b200 as q
code = q.Code.qc(4, 5, 10, 61, 9, 49)
dec = q.Decoder(code, 0, 1 << 20)
for W in (15, 30, 45):
    dec.get_statistics_weightw(W, 200000, 0.02, 100, 1234)
    t0 = time.time(); k = dec.get_statistics_weightw(W, 1000000, 0.02, 100, 1234)["counters"]; dt = time.time() - t0
    print("W=%d: %.3f s  %.2f M frames/s  corrected %d" % (W, dt, 1.0 / dt, int(k[3])))
