"""Launch shape the heuristic picks (bp_configure) and the BP kernel rate for the three reference code sizes."""
import sys
sys.path.insert(0, ".")
import qec_ldpc_b200 as q
import numpy as np
for prm, p, it, nf in [((3,3,6,7,2,3), 0.05, 20, 400000), ((4,5,10,61,9,49), 0.05, 50, 400000), ((4,4,8,509,208,2), 0.03, 50, 40000)]:
    code = q.Code.qc(*prm); dec = q.Decoder(code, 0, nf); dec.enable_timing(True)
    print(prm, dec.launch_info(0), dec.launch_info(1))
    dec.get_statistics_depolarizing(1, 0, nf, p, it); dec.get_timing(reset=True)
    k = dec.get_statistics_depolarizing(1, 0, nf, p, it)["counters"]; ms, _ = dec.get_timing()
    for s, nm in ((0, "bp_x"), (1, "bp_z")):
        print("   %s %.3f ms  %.3e edge-updates/s" % (nm, ms[nm], int(k[9+s]) * code.E[s] / ms[nm] * 1e3))
