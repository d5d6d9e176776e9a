"""HBM-resident path (both sides interleaved): fraction of the HBM roofline against the number of frame slots in flight.
Usage: python tools/global_slots_sweep.py C2|C5 [frames] [hbm_gbs]"""
import sys
sys.path.insert(0, ".")
import qec_ldpc_b200 as q

which = sys.argv[1] if len(sys.argv) > 1 else "C2"
prm, p, maxit, dflt = {"C2": ((4, 5, 10, 61, 9, 49), 0.05, 50, 1_000_000), "C5": ((4, 4, 8, 509, 208, 2), 0.03, 30, 100_000)}[which]
n = int(sys.argv[2]) if len(sys.argv) > 2 else dflt
peak = float(sys.argv[3]) if len(sys.argv) > 3 else 6549.1
code = q.Code.qc(*prm)
dec = q.Decoder(code, 0, n)
dec.enable_timing(True)
for slots in [int(s) for s in (sys.argv[4].split(",") if len(sys.argv) > 4 else "16384,32768,49152,65536,98304,131072".split(","))]:
    for side in (0, 1):
        dec.configure(side, -1, slots, 0)
    dec.get_statistics_depolarizing(1, 0, n, p, maxit)
    dec.get_timing(reset=True)
    k = dec.get_statistics_depolarizing(1, 0, n, p, maxit)["counters"]
    ms, _ = dec.get_timing(reset=True)
    t = ms["bp_x"] + ms["bp_z"]
    eu = int(k[9]) * code.E[0] + int(k[10]) * code.E[1]
    print("%s slots %7d : %8.2f ms  %.3e edge-updates/s  %.3f of %.0f GB/s" % (which, slots, t, eu / t * 1e3, 16 * eu / t * 1e3 / 1e9 / peak, peak), flush=True)
