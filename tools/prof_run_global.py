import sys
sys.path.insert(0, ".")
import qec_ldpc_b200 as q
code = q.Code.qc(4, 5, 10, 61, 9, 49)
n = 400000
dec = q.Decoder(code, 0, n)
for side in (0, 1):
    dec.configure(side, -1, 0, 0)
for _ in range(2):
    k = dec.get_statistics_depolarizing(1, 0, n, 0.05, 50)["counters"]
print(k)
