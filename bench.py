#!/usr/bin/env python
"""Benchmark of the decode hot path: Monte-Carlo BP decoding of quasi-cyclic quantum CSS LDPC codes under
depolarizing noise.

  python bench.py [--gpus N] [--steps K] [--warmup W] [--impl native|reference] [--config C1|C2|C3|C4|C5|C5g|C2g]

A "step" is one pass of the hot path over one batch of synthetic input -- Philox error generation + syndromes (one
kernel), BP on the X and Z Tanner graphs with fused hard decision / syndrome check / convergence test, and the
statistics reduction, all on device.  --config picks the BASELINE.json configuration (default C2 = configs[1], the one
the metric is quoted on):

  C1  J3K3L6P7 s2 t3      p=0.05  20 iterations     10,000 frames per GPU per step (the reference's CPU-sized case)
  C2  J4K5L10P61 s9 t49   p=0.05  50 iterations  1,000,000 frames per GPU per step
  C3  code610.txt (the same code, read from its 4-line text file), FER-vs-p sweep p = 0.01 .. 0.10 (10 points),
      100,000 frames per point per GPU per step, 50 iterations; every rank decodes its share of every point
  C4  J4K5L10P61          p=0.01 200 iterations  1,000,000 frames (early-exit divergence)
  C5  J4K4L8P509 s208 t2  p=0.03  50 iterations    100,000 frames (n=4072; default dispatch: shared-memory tile)
  C5g / C2g               the same through the HBM-resident kernel family (bp_global.cu)

Weak scaling: every rank decodes its own frame range per step (global frame ids, disjoint Philox streams); the only
collective is one NCCL all-reduce of the counter vector.

Printed JSON (rank 0, one line): the driver contract + "roofline" (the memory tier that holds the BP message state of
this configuration: shared memory for the tile kernel, HBM for the HBM-resident family), "cpu_baseline", "e2e",
"clocks", "gpu_launches" (see DESIGN.md section 6).
"""
import argparse
import json
import os
import sys
import tempfile
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

SEED = 0x5EED_B200
BYTES_PER_EDGE_UPDATE = 16  # fp32 flooding BP: 2 reads + 2 writes of the edge message per iteration
SWEEP_P = [round(0.01 * k, 2) for k in range(1, 11)]

CONFIGS = {
    "C1": dict(code=(3, 3, 6, 7, 2, 3), golden="C1", name="J3K3L6P7 s2 t3 (n=42)", segments=[(0.05, 10_000)], maxit=20),
    "C2": dict(code=(4, 5, 10, 61, 9, 49), golden="C2", name="J4K5L10P61 s9 t49 (n=610)", segments=[(0.05, 1_000_000)],
               maxit=50),
    "C3": dict(code=(4, 5, 10, 61, 9, 49), golden="C2", name="code610.txt = J4K5L10P61 s9 t49 (n=610), read from file",
               segments=[(p, 100_000) for p in SWEEP_P], maxit=50, from_file=True),
    "C4": dict(code=(4, 5, 10, 61, 9, 49), golden="C2", name="J4K5L10P61 s9 t49 (n=610)", segments=[(0.01, 1_000_000)],
               maxit=200),
    "C5": dict(code=(4, 4, 8, 509, 208, 2), golden=None, name="J4K4L8P509 s208 t2 (n=4072)", segments=[(0.03, 100_000)],
               maxit=50),
}
CONFIGS["C5g"] = dict(CONFIGS["C5"], force_global=True)
CONFIGS["C2g"] = dict(CONFIGS["C2"], force_global=True)


def config_of(args):
    cfg = dict(CONFIGS[args.config])
    if args.frames:  # frames per GPU per step, spread over the segments in proportion
        tot = sum(f for _, f in cfg["segments"])
        cfg["segments"] = [(p, max(1, f * args.frames // tot)) for p, f in cfg["segments"]]
    cfg["frames"] = sum(f for _, f in cfg["segments"])
    return cfg


def metric_name(args, cfg):
    ps = cfg["segments"]
    noise = "depolarizing p=%g" % ps[0][0] if len(ps) == 1 else "depolarizing p=%g..%g (%d points)" % (ps[0][0], ps[-1][0], len(ps))
    return "decoded frames/s, %s %s, %d BP iterations" % (cfg["name"].split()[0], noise, cfg["maxit"])


def workload_config(args, cfg, n_gpus):
    F = cfg["frames"]
    ps = cfg["segments"]
    noise = "depolarizing p=%g" % ps[0][0] if len(ps) == 1 else \
        "FER-vs-p sweep p=%s, %d frames per point" % (",".join("%g" % p for p, _ in ps), ps[0][1])
    state = "HBM-resident BP message state (bp_global.cu)" if cfg.get("force_global") else \
        "BP message state is shared-memory resident by design"
    return {"workload": "%s %s, %s, %d iterations, %d frames per GPU per step" % (args.config, cfg["name"], noise,
                                                                                   cfg["maxit"], F),
            "config": args.config, "code": cfg["name"], "p": [p for p, _ in ps] if len(ps) > 1 else ps[0][0],
            "max_iterations": cfg["maxit"], "frames_per_gpu_per_step": F, "global_frames_per_step": F * n_gpus,
            "parallelism": "frames sharded x%d, no data-path collective" % n_gpus,
            "l2": "per-step frame buffers (errors, syndromes, decisions: ~%d MB) %s the 126 MB L2; %s"
                  % (F * 0.0004 * (cfg["code"][2] * cfg["code"][3]) / 610, "exceed" if F >= 400_000 else "are re-written every step; compare with", state)}


# --------------------------------------------------------------------------------------------------------------------
# clocks
# --------------------------------------------------------------------------------------------------------------------
class ClockSampler(threading.Thread):
    REASONS = {0x4: "sw_power_cap", 0x8: "hw_slowdown", 0x20: "sw_thermal_slowdown", 0x40: "hw_thermal_slowdown",
               0x80: "hw_power_brake_slowdown", 0x2: "applications_clocks_setting", 0x10: "sync_boost"}

    def __init__(self, device_index):
        super().__init__(daemon=True)
        self.stop_flag = threading.Event()
        self.sm, self.reasons, self.max_mhz, self.ok = [], set(), None, False
        try:
            import pynvml
            import torch
            pynvml.nvmlInit()
            uuid = str(torch.cuda.get_device_properties(device_index).uuid)
            if not uuid.startswith("GPU-"):
                uuid = "GPU-" + uuid
            try:
                self.h = pynvml.nvmlDeviceGetHandleByUUID(uuid.encode())
            except Exception:
                self.h = pynvml.nvmlDeviceGetHandleByIndex(device_index)
            self.nv = pynvml
            self.max_mhz = int(pynvml.nvmlDeviceGetMaxClockInfo(self.h, pynvml.NVML_CLOCK_SM))
            self.ok = True
        except Exception:
            self.ok = False

    def run(self):
        if not self.ok:
            return
        while not self.stop_flag.is_set():
            try:
                self.sm.append(int(self.nv.nvmlDeviceGetClockInfo(self.h, self.nv.NVML_CLOCK_SM)))
                try:
                    r = int(self.nv.nvmlDeviceGetCurrentClocksEventReasons(self.h))
                except Exception:
                    r = int(self.nv.nvmlDeviceGetCurrentClocksThrottleReasons(self.h))
                for bit, name in self.REASONS.items():
                    if r & bit:
                        self.reasons.add(name)
            except Exception:
                pass
            self.stop_flag.wait(0.1)

    def result(self):
        self.stop_flag.set()
        if self.is_alive():
            self.join(1.0)
        if not self.sm:
            return {"sm_mhz": None, "sm_max_mhz": self.max_mhz, "reasons": ["unavailable"], "samples": 0}
        return {"sm_mhz": float(np.median(self.sm)), "sm_max_mhz": self.max_mhz, "reasons": sorted(self.reasons),
                "samples": len(self.sm)}


# --------------------------------------------------------------------------------------------------------------------
# CPU reference (oracle/_ref = the unmodified reference; oracle port as fall-back).  Nothing in this section touches
# the product package: the code file the reference needs is written with numpy from the golden matrices.
# --------------------------------------------------------------------------------------------------------------------
def golden_matrix(code, key):
    g = np.load(os.path.join(ROOT, "tests", "golden", "codes.npz"))
    shp = g["%s_%s_shape" % (code, key)]
    return np.unpackbits(g["%s_%s" % (code, key)], axis=1)[:, :shp[1]].astype(np.int32)


def write_code_file(path, params, pcmX, pcmZ, imp):
    """The reference's 4-line text format (Quantum_LDPC_Code.h:26-80): `J K L P sigma tau`, then pcmX, pcmZ, iMinusP
    as tab-separated 0/1, no trailing newline."""
    def row(m):
        return b"\t".join([b"0", b"1"][int(v)] for v in np.asarray(m, np.uint8).ravel())
    with open(path, "wb") as f:
        f.write(("\t".join(str(int(v)) for v in params) + "\n").encode())
        f.write(row(pcmX) + b"\n")
        f.write(row(pcmZ) + b"\n")
        f.write(row(imp))


def golden_code_file(golden):
    g = np.load(os.path.join(ROOT, "tests", "golden", "codes.npz"))
    path = os.path.join(tempfile.mkdtemp(prefix="qldpc_"), "code.txt")
    write_code_file(path, g[golden + "_params"], golden_matrix(golden, "pcmX"), golden_matrix(golden, "pcmZ"),
                    golden_matrix(golden, "iMinusP"))
    return path


class CpuReference:
    """Times the reference's own CPU implementation of the path (Decode + GetSyndrome + CheckLogicalError per frame,
    one DecoderCPU per OpenMP thread, DecoderCPU.h:419-431) on the host cores, on the same Philox patterns."""

    def __init__(self, cfg):
        from oracle.pyoracle import Oracle, Reference, build
        build(ref=True)
        self.cfg = cfg
        self.oracle = Oracle()
        self.oc = self.oracle.code_qc(*cfg["code"])
        self.kind = "port"
        self.rc = None
        self.note = ""
        if cfg["golden"]:
            self.oc.set_logical(golden_matrix(cfg["golden"], "iMinusP"))
            if Reference.available():
                # the reference loads codes from its 4-line text files only (Quantum_LDPC_Code.h:26-80)
                self.rc = Reference().code_from_file(golden_code_file(cfg["golden"]))
                self.kind = "reference"
        else:
            self.note = ("; no code file exists for this code (the reference needs one with iMinusP), so the oracle "
                         "port runs it, without the logical check")
        # all host cores this process may use (torchrun exports OMP_NUM_THREADS=1; the harnesses set the count themselves)
        self.cores = len(os.sched_getaffinity(0)) if hasattr(os, "sched_getaffinity") else (os.cpu_count() or 1)

    def patterns(self, first, n, p):
        return self.oc.depolarizing_bulk(SEED, first, n, p)

    def run(self, x, z, p):
        """Returns (seconds, corrected)."""
        if self.rc is not None:
            out = self.rc.run_frames(x, z, p, self.cfg["maxit"], self.cores)
            return out["seconds"], out["counters"]["corrected"]
        out = self.oc.run_frames(x, z, p, self.cfg["maxit"], self.cores)
        return out["seconds"], int(out["counters"][3])

    def sample(self, first, total):
        """`total` frames spread over the configuration's noise points in proportion: (seconds, frames, corrected)."""
        segs = self.cfg["segments"]
        tot = sum(f for _, f in segs)
        sec = frames = corrected = 0
        for p, f in segs:
            cnt = max(1, total * f // tot)
            x, z = self.patterns(first, cnt, p)
            s, c = self.run(x, z, p)
            sec, frames, corrected, first = sec + s, frames + cnt, corrected + c, first + cnt
        return sec, frames, corrected

    def calibrate(self, target_s):
        n0 = max(16 * len(self.cfg["segments"]), 4 * self.cores)
        sec, frames, _ = self.sample(0, n0)
        rate = frames / max(sec, 1e-6)
        return int(min(400_000, max(n0, rate * target_s)))


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    cfg = config_of(args)
    ref = CpuReference(cfg)
    sample = ref.calibrate(3.0)
    times, frames = [], 0
    for s in range(args.warmup + args.steps):
        sec, cnt, _ = ref.sample(s * sample, sample)
        if s >= args.warmup:
            times.append(sec)
            frames += cnt
    total = float(np.sum(times))
    value = frames / total
    desc = "%d frames per step of the same Philox stream (frame ids from 0), %d OpenMP threads, g++ -O2 -fopenmp%s" % (
        frames // args.steps, ref.cores, ref.note)
    line = {"impl": "reference", "metric": metric_name(args, cfg), "value": value, "unit": "frames/s", "n_gpus": args.gpus,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * total / args.steps,
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": workload_config(args, cfg, args.gpus),
            "cpu_baseline": {"value": value, "unit": "frames/s", "cores": ref.cores, "kind": ref.kind, "sample": desc},
            "e2e": {"value": value, "unit": "frames/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0}
    print(json.dumps(line), flush=True)


# --------------------------------------------------------------------------------------------------------------------
# native arm
# --------------------------------------------------------------------------------------------------------------------
def run_native(args):
    import torch
    import torch.distributed as dist
    import qec_ldpc_b200 as q

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device; the decode path has no CPU fallback")
    torch.cuda.set_device(local)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    n_gpus = world
    cfg = config_of(args)
    F, segs, maxit = cfg["frames"], cfg["segments"], cfg["maxit"]

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    if cfg.get("from_file"):
        code = q.Code.from_file(golden_code_file(cfg["golden"]))
    else:
        code = q.Code.qc(*cfg["code"])
    dec = q.Decoder(code, local, max(f for _, f in segs))
    if cfg.get("force_global"):
        for side in (0, 1):
            dec.configure(side, -1, 0, 0)
    stream = torch.cuda.Stream()
    dec.set_stream(stream.cuda_stream)
    info = [dec.launch_info(0), dec.launch_info(1)]
    global_path = info[0]["vec"] < 0 or info[1]["vec"] < 0

    def first_frame(step):  # global frame ids: disjoint Philox streams per (step, rank)
        return (step * world + rank) * F

    def device_step(step):
        """One step from scalars: errors generated on the device.  Returns the per-segment counters."""
        out, first = [], first_frame(step)
        for p, f in segs:
            out.append(dec.get_statistics_depolarizing(SEED, first, f, p, maxit)["counters"])
            first += f
        return out

    # ---- device-resident measurement ("value") ---------------------------------------------------------------
    for s in range(args.warmup):
        device_step(10_000 + s)
    dec.get_timing(reset=True)
    dec.enable_timing(True)
    sampler = ClockSampler(local)
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    seg_counters = np.zeros((len(segs), q.NUM_COUNTERS), np.uint64)
    barrier()
    sampler.start()
    ev0.record(stream)
    for s in range(args.steps):
        seg_counters += np.stack(device_step(s))
    ev1.record(stream)
    barrier()
    clocks = sampler.result()
    ms_total = ev0.elapsed_time(ev1)
    kms, klaunch = dec.get_timing(reset=True)
    dec.enable_timing(False)
    counters = seg_counters.sum(axis=0)

    # the only data collective: one NCCL all-reduce (sum) of the counter vector; the time is max-reduced over ranks
    from qec_ldpc_b200.sharding import allreduce_counters, allreduce_max
    ms_total = allreduce_max(ms_total, "cuda")
    gseg = allreduce_counters(seg_counters.ravel(), "cuda").reshape(seg_counters.shape)
    gc = gseg.sum(axis=0)

    # ---- end to end through the C ABI with HOST buffers ("e2e") ----------------------------------------------
    # DecoderGPU::GetStats(..., xErrors, zErrors) (DecoderGPU.h:193): pre-generated patterns in host memory, in the
    # reference's layout (one int per qubit, frame-major); host marshalling, H2D copies, decode and the D2H counter
    # read are all inside the timed region.  The call has two ways to marshal the rows: pack them to bits with host
    # threads before the copy, or copy the raw rows and pack on the device.  BOTH are timed every time, at every rank
    # count (host_packed: min(16, cores / local ranks) threads; raw_rows_over_link: threads = 0); `value` is the one the
    # library picks by default on this box (packing needs >= 10 threads per rank to beat the raw copy of int32 rows),
    # named in `method`.
    n = code.n
    code_nw = (n + 31) // 32
    xh = torch.empty((F, n), dtype=torch.int32, pin_memory=True)
    zh = torch.empty((F, n), dtype=torch.int32, pin_memory=True)
    sl, off = 100_000, 0
    for p, f in segs:
        for o in range(0, f, sl):
            cnt = min(sl, f - o)
            x, z, _, _ = dec.debug_generate(SEED, first_frame(0) + off + o, cnt, p)
            xh[off + o:off + o + cnt] = torch.from_numpy(x)
            zh[off + o:off + o + cnt] = torch.from_numpy(z)
        off += f
    host_threads = q.default_host_threads()
    in_use = dec.host_threads_in_use(4)

    def host_step(xt, zt, elem):
        tot, off = np.zeros(q.NUM_COUNTERS, np.uint64), 0
        for p, f in segs:
            tot += dec.get_stats_from_errors_ptr(xt.data_ptr() + off * n * elem, zt.data_ptr() + off * n * elem, f, p,
                                                 maxit, elem=elem)
            off += f
        return tot

    def time_host(xt, zt, elem, threads):
        dec.set_host_threads(threads)
        out = None
        for s in range(max(1, min(args.warmup, 2))):
            host_step(xt, zt, elem)
        barrier()
        t0 = time.perf_counter()
        for s in range(args.steps):
            out = host_step(xt, zt, elem)
        torch.cuda.synchronize()
        return allreduce_max(time.perf_counter() - t0, "cuda"), out

    e2e_raw_s, raw_counters = time_host(xh, zh, 4, 0)
    e2e_packed_s, e2e_counters = time_host(xh, zh, 4, host_threads)
    assert np.array_equal(raw_counters, e2e_counters)
    e2e_s = e2e_packed_s if in_use > 0 else e2e_raw_s
    # same patterns as step 0 of the device-resident run => same counters (checked on every rank)
    chk = np.stack(device_step(0)).sum(axis=0)
    assert np.array_equal(chk, e2e_counters), "host-buffer path and device-generated path disagree"
    # host-memory roofline of the int32 layout: every rank streams its 2 x F x n x 4 bytes through the host memory bus
    # once per step; the denominator is the streaming-read bandwidth the SAME threads reach on the SAME buffers with all
    # ranks reading at once (qldpc_debug_host_read_gbs)
    barrier()
    host_gbs = q.host_read_gbs(xh.data_ptr(), F * n * 4, host_threads, 3)
    host_gbs_all = host_gbs
    if world > 1:
        t = torch.tensor([host_gbs], dtype=torch.float64, device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.SUM)
        host_gbs_all = float(t.item())

    # two more end-to-end views, reported next to the primary one: (a) the same call with one BYTE per qubit
    # (qldpc_get_stats_from_errors_u8: a quarter of the host bytes), (b) the Monte-Carlo call itself,
    # qldpc_get_statistics_depolarizing (what GetStatistics(W, COUNT, p, MAXIT) is in the reference's driver,
    # main.cu:101): scalars in, counters out, errors generated on the device -- timed on the host clock.
    xb, zb = xh.to(torch.uint8).pin_memory(), zh.to(torch.uint8).pin_memory()
    e2e_u8_s, u8_counters = time_host(xb, zb, 1, -1)  # library default for byte rows
    in_use8 = dec.host_threads_in_use(1)
    assert np.array_equal(chk, u8_counters)
    barrier()
    t0 = time.perf_counter()
    for s in range(args.steps):
        device_step(s)
    torch.cuda.synchronize()
    e2e_gs_s = allreduce_max(time.perf_counter() - t0, "cuda")
    del xb, zb

    if rank == 0:
        frames = int(gc[0])
        eu = int(gc[9]) * code.EX + int(gc[10]) * code.EZ
        secs = ms_total * 1e-3
        value = frames / secs
        # roofline of the dominant kernel (the BP launches of both sides), rank 0's launches
        k_eu = int(counters[9]) * code.EX + int(counters[10]) * code.EZ
        bp_ms = kms["bp_x"] + kms["bp_z"]
        bp_launches = klaunch["bp_x"] + klaunch["bp_z"]
        achieved = k_eu * BYTES_PER_EDGE_UPDATE / (bp_ms * 1e-3) / 1e9
        peaks_path = os.path.join(ROOT, "MEASURED_PEAKS.json")
        if os.path.exists(peaks_path):
            hbm_peak, hbm_src = float(json.load(open(peaks_path))["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
        else:
            hbm_peak, hbm_src = 6650.0, "fallback (B200_PROFILING.md)"
        traffic = None
        tpath = os.path.join(ROOT, "profiles", "traffic.json")
        if os.path.exists(tpath):
            traffic = json.load(open(tpath)).get("dram_bytes_per_launch_by_config", {}).get(args.config)
        sm_mhz = clocks["sm_mhz"] or 1965.0
        smem_nominal = 148 * 128 * sm_mhz * 1e6 / 1e9  # GB/s at the SM clock sampled under load
        smem_peak, smem_src = smem_nominal, "nominal 148 SM x 128 B/clk x %.0f MHz (sampled under load)" % sm_mhz
        spath = os.path.join(ROOT, "profiles", "smem_peak.json")
        if os.path.exists(spath):  # measured by tools/micro/smem_bw.cu (LDS.128 + STS.128, the kernel's 1:1 mix)
            smem_peak = float(json.load(open(spath))["roofline_denominator_gbs"])
            smem_src = "measured shared-memory copy bandwidth (profiles/smem_peak.json, tools/micro/smem_bw.cu); " \
                       "nominal %.0f GB/s" % smem_nominal
        step_kernel_ms = sum(kms.values())
        common = {"achieved": achieved, "unit": "GB/s", "traffic": traffic,
                  "algorithmic_bytes_per_edge_update": BYTES_PER_EDGE_UPDATE,
                  "edge_updates_per_launch": k_eu / max(bp_launches, 1), "avg_launch_ms": bp_ms / max(bp_launches, 1),
                  "kernel_share_of_step": bp_ms / max(step_kernel_ms, 1e-9)}
        if global_path:
            roofline = dict(common, bound="hbm", peak=hbm_peak, frac=achieved / hbm_peak, peak_source=hbm_src,
                            kernel="g_check / g_var (bp_global.cu), X + Z sides; launch = one side's whole run",
                            note="messages live in HBM (slot-innermost arrays): 16 B per edge-update cross the HBM bus")
            views = {}
        else:
            roofline = dict(common, bound="smem", peak=smem_peak, frac=achieved / smem_peak, peak_source=smem_src,
                            kernel="bp_tile_kernel (X + Z launches)",
                            note="the BP message state is shared-memory resident for the whole decode (SURVEY 8(d): the "
                                 "relevant tier for this configuration); `traffic` is the DRAM traffic of a launch "
                                 "(syndromes in, decisions out), which shows that the state never touches HBM")
            views = {"hbm_view": {"bound": "hbm", "achieved": achieved, "peak": hbm_peak, "unit": "GB/s",
                                  "frac": achieved / hbm_peak, "peak_source": hbm_src,
                                  "note": "the same algorithmic bytes against the HBM copy peak: > 1 because the state is "
                                          "on chip; what an HBM-resident design would be bound by"}}
        apath = os.path.join(ROOT, "profiles", "r2", "arith_ceiling.json")
        if os.path.exists(apath) and not global_path and args.config in ("C2", "C3", "C4"):
            # measured arithmetic-only ceiling of the reference's operation order (tools/micro/arith_ceiling.cu)
            ce = {r["ceiling"]: r["edge_updates_per_s"] for r in json.load(open(apath))["streams"]
                  if "ceiling" in r and r["warps_per_sm"] == 28}
            cx = ce.get("check dc=10 + var dv=4 (scalar den adds)")
            cz = ce.get("check dc=10 + var dv=5 (scalar den adds)")
            if cx and cz:
                ex, ez = int(counters[9]) * code.EX, int(counters[10]) * code.EZ
                ceil_eu = (ex + ez) / (ex / cx + ez / cz)
                views["arithmetic_ceiling"] = {
                    "bound": "fp32 arithmetic of the reference's operation order, no memory (measured microbenchmark)",
                    "achieved": k_eu / (bp_ms * 1e-3), "peak": ceil_eu, "unit": "edge-updates/s",
                    "frac": k_eu / (bp_ms * 1e-3) / ceil_eu,
                    "as_smem_roofline_frac": ceil_eu * BYTES_PER_EDGE_UPDATE / 1e9 / smem_peak,
                    "peak_source": "profiles/r2/arith_ceiling.json (tools/micro/arith_ceiling.cu: the exact FMUL2 / FFMA2 / "
                                   "FADD / MUFU stream on registers, 28 warps per SM)"}
        host_bytes = 2 * F * n * 4
        line = {
            "metric": metric_name(args, cfg), "value": value, "unit": "frames/s", "n_gpus": n_gpus, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": ms_total / args.steps, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "f32", "data": "synthetic", "config": workload_config(args, cfg, n_gpus),
            "edge_updates_per_s": eu / secs,
            "mean_iterations": {"x": int(gc[9]) / frames, "z": int(gc[10]) / frames},
            "frame_error_rate": 1.0 - int(gc[3]) / frames,
            "counters": dict(zip(q.COUNTER_NAMES, [int(v) for v in gc])),
            "roofline": roofline,
            "kernel_ms_per_step": {k: v / args.steps for k, v in kms.items()},
            "launch": {"x": info[0], "z": info[1]},
            "e2e": {"value": F * n_gpus * args.steps / e2e_s, "unit": "frames/s",
                    "h2d_bytes_per_step": 2 * F * code_nw * 4 if in_use > 0 else host_bytes,
                    "d2h_bytes_per_step": q.NUM_COUNTERS * 8 * len(segs),
                    "host_buffer_bytes_per_step": host_bytes, "host_threads": in_use,
                    "method": "host_packed (%d threads per rank)" % in_use if in_use > 0 else "raw_rows_over_link",
                    "api": "qldpc_get_stats_from_errors_i32 (DecoderGPU::GetStats layout, pinned host int32)",
                    "note": "the reference's layout spends one int32 per qubit (%d B per frame); `value` is the library's "
                            "default marshalling on this box (`method`); host_packed and raw_rows_over_link are both timed "
                            "at every rank count, so either series can be followed on its own; u8_patterns and "
                            "device_generated show the other two entry points" % (2 * n * 4),
                    "ms_per_step": 1e3 * e2e_s / args.steps,
                    "vs_device_value": (F * n_gpus * args.steps / e2e_s) / value,
                    "host_mem_roofline": {"bound": "host memory read bandwidth", "unit": "GB/s",
                                          "bytes_per_step_all_ranks": host_bytes * n_gpus,
                                          "achieved": host_bytes * n_gpus * args.steps / e2e_s / 1e9,
                                          "peak": host_gbs_all,
                                          "frac": host_bytes * n_gpus * args.steps / e2e_s / 1e9 / host_gbs_all,
                                          "peak_source": "measured: streaming read of the same pinned buffers by the same "
                                                         "%d threads per rank, all %d ranks at once "
                                                         "(qldpc_debug_host_read_gbs)" % (host_threads, n_gpus)},
                    "host_packed": {"value": F * n_gpus * args.steps / e2e_packed_s, "unit": "frames/s",
                                    "h2d_bytes_per_step": 2 * F * code_nw * 4, "host_threads": host_threads},
                    "raw_rows_over_link": {"value": F * n_gpus * args.steps / e2e_raw_s, "unit": "frames/s",
                                           "h2d_bytes_per_step": host_bytes, "host_threads": 0},
                    "u8_patterns": {"value": F * n_gpus * args.steps / e2e_u8_s, "unit": "frames/s",
                                    "h2d_bytes_per_step": 2 * F * code_nw * 4 if in_use8 > 0 else 2 * F * n,
                                    "host_buffer_bytes_per_step": 2 * F * n, "host_threads": in_use8,
                                    "method": "host_packed (%d threads per rank)" % in_use8 if in_use8 > 0
                                              else "raw_rows_over_link",
                                    "vs_device_value": (F * n_gpus * args.steps / e2e_u8_s) / value,
                                    "api": "qldpc_get_stats_from_errors_u8"},
                    "device_generated": {"value": F * n_gpus * args.steps / e2e_gs_s, "unit": "frames/s",
                                         "h2d_bytes_per_step": 0, "d2h_bytes_per_step": q.NUM_COUNTERS * 8 * len(segs),
                                         "api": "qldpc_get_statistics_depolarizing (host wall clock)"}},
            "gpu_launches": int(sum(klaunch.values())),
            "clocks": clocks,
        }
        line.update(views)
        if len(segs) > 1:  # FER-vs-p sweep: per-point rates with Wilson 95% intervals
            pts = []
            for (p, _), k in zip(segs, gseg):
                nfr, bad = int(k[0]), int(k[0]) - int(k[3])
                ph, z = bad / nfr, 1.959964
                den = 1 + z * z / nfr
                ctr, half = (ph + z * z / (2 * nfr)) / den, z * np.sqrt(ph * (1 - ph) / nfr + z * z / (4.0 * nfr * nfr)) / den
                pts.append({"p": p, "frames": nfr, "frame_errors": bad, "fer": ph, "wilson95": [ctr - half, ctr + half],
                            "logical": int(k[6]), "mean_iterations": [int(k[9]) / nfr, int(k[10]) / nfr]})
            line["sweep"] = pts
        if n_gpus == 1 and not args.no_cpu:
            ref = CpuReference(cfg)
            sample = ref.calibrate(args.cpu_seconds)
            sec, cnt, corrected = ref.sample(0, sample)
            line["cpu_baseline"] = {
                "value": cnt / sec, "unit": "frames/s", "cores": ref.cores, "kind": ref.kind,
                "sample": "first %d frames of the same Philox stream (seed %d), %.1f s wall, corrected %d%s"
                          % (cnt, SEED, sec, corrected, ref.note)}
        else:
            line["cpu_baseline"] = None
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="native", choices=["native", "reference"])
    ap.add_argument("--config", default="C2", choices=sorted(CONFIGS))
    ap.add_argument("--frames", type=int, default=0, help="frames per GPU per step (0 = the configuration's own)")
    ap.add_argument("--no-cpu", action="store_true", help="skip the cpu_baseline leg")
    ap.add_argument("--cpu-seconds", type=float, default=12.0)
    args = ap.parse_args()
    if args.impl == "reference":
        run_reference(args)
    else:
        run_native(args)


if __name__ == "__main__":
    main()
