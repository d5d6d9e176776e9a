"""Launch-shape sweep for the BP tile kernel (p=0.05, 50 iterations): tile width x threads per CTA.
Usage: python tools/tune.py [frames] [J,K,L,P,sigma,tau]"""
import sys
sys.path.insert(0, ".")
import qec_ldpc_b200 as q

n = int(sys.argv[1]) if len(sys.argv) > 1 else 400_000
prm = [int(v) for v in sys.argv[2].split(",")] if len(sys.argv) > 2 else [4, 5, 10, 61, 9, 49]
code = q.Code.qc(*prm)
dec = q.Decoder(code, 0, n)
dec.enable_timing(True)
best = {}
for side in (0, 1):
    for vec in (4, 2, 1):
        for thr in (32, 64, 96, 128, 160, 192, 224, 256, 384, 512):
            try:
                dec.configure(side, vec, thr, 0)
            except q.QldpcError:
                continue
            info = dec.launch_info(side)
            dec.get_statistics_depolarizing(1, 0, n, 0.05, 50)
            dec.get_timing(reset=True)
            k = dec.get_statistics_depolarizing(1, 0, n, 0.05, 50)["counters"]
            ms, _ = dec.get_timing(reset=True)
            t = ms["bp_x" if side == 0 else "bp_z"]
            eu = int(k[9 + side]) * code.E[side]
            print("side %d vec %d thr %3d ctas/SM %2d regs %3d smem %6d : %7.3f ms  %.3e edge-updates/s" % (
                side, vec, thr, info["ctas_per_sm"], info["regs"], info["smem"], t, eu / t * 1e3), flush=True)
            if side not in best or t < best[side][0]:
                best[side] = (t, vec, thr)
    dec.configure(side, 0, 0, 0)
print("best", best)
