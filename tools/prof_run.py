"""Small driver for ncu captures: one Monte-Carlo pass of J4K5L10P61 (p=0.05, 50 iterations) over N frames."""
import sys
sys.path.insert(0, ".")
import qec_ldpc_b200 as q

n = int(sys.argv[1]) if len(sys.argv) > 1 else 200_000
code = q.Code.qc(4, 5, 10, 61, 9, 49)
dec = q.Decoder(code, 0, n)
for _ in range(2):
    k = dec.get_statistics_depolarizing(1, 0, n, 0.05, 50)["counters"]
print(dict(zip(q.COUNTER_NAMES, [int(v) for v in k])))
