import sys
sys.path.insert(0, ".")
import numpy as np
import qec_ldpc_b200 as q
from oracle.pyoracle import Oracle
O = Oracle()
prm = (3, 3, 6, 7, 2, 3)
code = q.Code.qc(*prm); oc = O.code_qc(*prm)
dec = q.Decoder(code, 0, 4096)
p, maxit, nf = 0.05, 20, 8
_, _, sx, sz = dec.debug_generate(11, 0, nf, p)
for vec in (4, 2, 1):
    dec.configure(0, vec, 0, 0)
    qt, rt, it = dec.debug_bp_trace(0, sx, p, maxit, maxit)
    for f in range(nf):
        oit, _, _, oq, orr = oc.bp(0, sx[f], p, maxit, trace=maxit)
        for name, a, b in (("r", rt[f, :oit], orr[:oit]), ("q", qt[f, :oit], oq[:oit])):
            bad = np.argwhere((a.view(np.uint32) != b.view(np.uint32)) & ~(np.isnan(a) & np.isnan(b)))
            if len(bad):
                i, e = bad[0]
                print("vec", vec, "frame", f, name, "first mismatch iter", i, "edge", e, "(check", e // 6, "pos", e % 6, ") gpu", a[i, e], a[i, e].view(np.uint32), "cpu", b[i, e], b[i, e].view(np.uint32), "count", len(bad), "its", it[f], oit)
                break
        else:
            continue
        break
    else:
        print("vec", vec, "all frames identical")
