"""Summarises an .ncu-rep (raw + source pages) for the BP kernels: headline counters, opcode mix, stall mix.
Usage: python tools/ncu_summary.py gpurun_out/prof.ncu-rep [kernel-index]"""
import collections
import csv
import subprocess
import sys

rep = sys.argv[1]
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(raw.splitlines()))
hdr, units = rows[0], rows[1]
want = ["Kernel Name", "gpu__time_duration.sum", "smsp__inst_executed.sum", "sm__inst_executed.avg.per_cycle_elapsed",
        "smsp__issue_active.avg.pct_of_peak_sustained_active", "sm__warps_active.avg.pct_of_peak_sustained_active",
        "launch__registers_per_thread", "launch__grid_size", "launch__block_size",
        "launch__occupancy_limit_registers", "launch__occupancy_limit_shared_mem",
        "dram__bytes_read.sum", "dram__bytes_write.sum", "lts__t_sector_hit_rate.pct",
        "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum", "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum",
        "smsp__thread_inst_executed_per_inst_executed.ratio", "sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active",
        "sm__pipe_alu_cycles_active.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active",
        "sm__throughput.avg.pct_of_peak_sustained_elapsed", "l1tex__throughput.avg.pct_of_peak_sustained_elapsed",
        "smsp__cycles_active.avg", "sm__cycles_elapsed.avg", "sm__cycles_elapsed.avg.per_second"]
for w in want:
    if w in hdr:
        i = hdr.index(w)
        print("%-70s %-12s %s" % (w, units[i], [r[i] for r in rows[2:]]))
src = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
kern, cur = [], None
for r in csv.reader(src.splitlines()):
    if r and r[0] == "Kernel Name":
        cur = {"name": r[1], "rows": []}
        kern.append(cur)
    elif r and r[0] == "Address":
        cur["hdr"] = r
    elif cur is not None and len(r) > 5:
        cur["rows"].append(r)
for k in kern:
    h = k["hdr"]
    iI, iS, iSamp = h.index("Instructions Executed"), h.index("Source"), h.index("# Samples")
    iW, iWi = h.index("L1 Wavefronts Shared"), h.index("L1 Wavefronts Shared Ideal")
    tot = sum(int(r[iI]) for r in k["rows"])
    byop = collections.Counter()
    for r in k["rows"]:
        s = r[iS].strip()
        op = s.split()[1] if s.startswith("@") else s.split()[0]
        byop[op.split(".")[0]] += int(r[iI])
    print("\n" + k["name"], "warp-instructions", tot, "SASS lines", len(k["rows"]))
    print("  opcode mix: " + ", ".join("%s %.1f%%" % (o, 100 * c / tot) for o, c in byop.most_common(22)))
    w = sum(int(r[iW]) for r in k["rows"])
    wi = sum(int(r[iWi]) for r in k["rows"])
    print("  smem wavefronts %d (ideal %d, excess %.1f%%)" % (w, wi, 100.0 * (w - wi) / max(wi, 1)))
    st = [c for c in h if c.startswith("stall_") and "Not Issued" not in c]
    tots = {c: sum(int(r[h.index(c)]) for r in k["rows"]) for c in st}
    ts = sum(tots.values())
    print("  stall samples: " + ", ".join("%s %.1f%%" % (c[6:], 100 * v / ts) for c, v in sorted(tots.items(), key=lambda x: -x[1])[:9]))
    # region shares between barriers
    acc, last, regions = 0, 0, []
    for i, r in enumerate(k["rows"]):
        acc += int(r[iI])
        if "BAR.SYNC" in r[iS] or "EXIT" in r[iS] or "RET" in r[iS]:
            regions.append((i, r[iS].strip().split()[0], round(100.0 * (acc - last) / tot, 1)))
            last = acc
    print("  share of instructions between barriers (SASS line, marker, %):", [x for x in regions if x[2] >= 0.5])
