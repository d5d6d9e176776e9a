"""Occupancy sensitivity of the BP tile kernel: the default shape of each side with the resident CTAs per SM capped.
Usage: python tools/occ_sweep.py [frames]"""
import sys
sys.path.insert(0, ".")
import qec_ldpc_b200 as q

n = int(sys.argv[1]) if len(sys.argv) > 1 else 400_000
code = q.Code.qc(4, 5, 10, 61, 9, 49)
dec = q.Decoder(code, 0, n)
dec.enable_timing(True)
for side in (0, 1):
    for vec, thr in ((2, 128), (2, 64), (4, 128)):
        for ctas in (8, 7, 6, 5, 4, 3, 2):
            try:
                dec.configure(side, vec, thr, ctas)
            except q.QldpcError:
                continue
            info = dec.launch_info(side)
            if info["ctas_per_sm"] != ctas:
                continue
            dec.get_statistics_depolarizing(1, 0, n, 0.05, 50)
            dec.get_timing(reset=True)
            k = dec.get_statistics_depolarizing(1, 0, n, 0.05, 50)["counters"]
            ms, _ = dec.get_timing(reset=True)
            t = ms["bp_x" if side == 0 else "bp_z"]
            eu = int(k[9 + side]) * code.E[side]
            print("side %d vec %d thr %3d ctas/SM %2d (%2d warps) : %7.3f ms  %.3e edge-updates/s" % (
                side, vec, thr, ctas, ctas * thr // 32, t, eu / t * 1e3), flush=True)
    dec.configure(side, 0, 0, 0)
