#!/usr/bin/env python
"""Benchmark of the decode hot path: Monte-Carlo BP decoding of J4K5L10P61 (n=610) under depolarizing noise.

  python bench.py [--gpus N] [--steps K] [--warmup W] [--impl native|reference]

A "step" is one pass of the hot path over one batch of synthetic input: 1,000,000 frames of depolarizing(p=0.05)
noise, 50 BP iterations (BASELINE.json configs[1]) -- Philox error generation, syndromes, BP on the X and Z Tanner
graphs with fused hard decision / syndrome check / convergence test, and the statistics reduction, all on device.
Weak scaling: every rank decodes its own 1M-frame range per step (global frame ids, disjoint Philox streams); the
only collective is one NCCL all-reduce of the counter vector.

Printed JSON (rank 0, one line): the driver contract + "roofline", "cpu_baseline", "e2e", "clocks",
"gpu_launches" (see DESIGN.md section 6).
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

CODE = (4, 5, 10, 61, 9, 49)
P, MAXIT = 0.05, 50
FRAMES_PER_STEP = 1_000_000
SEED = 0x5EED_B200
BYTES_PER_EDGE_UPDATE = 16  # fp32 flooding BP: 2 reads + 2 writes of the edge message per iteration
METRIC = "decoded frames/s, J4K5L10P61 depolarizing p=0.05, 50 BP iterations"


def workload_config(n_gpus, frames):
    return {"workload": "J4K5L10P61 s9 t49 (n=610), depolarizing p=0.05, 50 iterations, %d frames per GPU per step" % frames,
            "code": "J4K5L10P61 s9 t49", "p": P, "max_iterations": MAXIT, "frames_per_gpu_per_step": frames,
            "global_frames_per_step": frames * n_gpus, "parallelism": "frames sharded x%d, no data-path collective" % n_gpus,
            "l2": "per-step frame buffers (errors, syndromes, decisions: ~400 MB) exceed the 126 MB L2; BP message "
                  "state is shared-memory resident by design"}


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
# CPU reference (oracle/_ref = the unmodified reference; oracle port as fall-back)
# --------------------------------------------------------------------------------------------------------------------
def golden_matrix(code, key):
    g = np.load(os.path.join(ROOT, "tests", "golden", "codes.npz"))
    shp = g["%s_%s_shape" % (code, key)]
    return np.unpackbits(g["%s_%s" % (code, key)], axis=1)[:, :shp[1]].astype(np.int32)


class CpuReference:
    """Times the reference's own CPU implementation of the path (Decode + GetSyndrome + CheckLogicalError per frame,
    one DecoderCPU per OpenMP thread, DecoderCPU.h:419-431) on the host cores, on the same Philox patterns."""

    def __init__(self):
        from oracle.pyoracle import Oracle, Reference, build
        build(ref=True)
        self.oracle = Oracle()
        self.oc = self.oracle.code_qc(*CODE)
        imp = golden_matrix("C2", "iMinusP")
        self.oc.set_logical(imp)
        self.kind = "port"
        self.rc = None
        if Reference.available():
            # the reference loads codes from its 4-line text files only (Quantum_LDPC_Code.h:26-80): write one
            import qec_ldpc_b200 as q
            code = q.Code.dense(*CODE, golden_matrix("C2", "pcmX"), golden_matrix("C2", "pcmZ"), imp)
            path = os.path.join(tempfile.mkdtemp(prefix="qldpc_"), "code610.txt")
            code.write_file(path)
            self.rc = Reference().code_from_file(path)
            self.kind = "reference"
        # all host cores this process may use (torchrun exports OMP_NUM_THREADS=1; the harnesses set the count themselves)
        self.cores = len(os.sched_getaffinity(0)) if hasattr(os, "sched_getaffinity") else (os.cpu_count() or 1)

    def patterns(self, first, n):
        return self.oc.depolarizing_bulk(SEED, first, n, P)

    def run(self, x, z):
        """Returns (seconds, corrected, edge_updates or None)."""
        if self.rc is not None:
            out = self.rc.run_frames(x, z, P, MAXIT, self.cores)
            return out["seconds"], out["counters"]["corrected"]
        out = self.oc.run_frames(x, z, P, MAXIT, self.cores)
        return out["seconds"], int(out["counters"][3])

    def calibrate(self, target_s):
        n0 = max(64, 16 * self.cores)
        x, z = self.patterns(0, n0)
        sec, _ = self.run(x, z)
        rate = n0 / max(sec, 1e-6)
        return int(min(400_000, max(n0, rate * target_s)))


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    ref = CpuReference()
    sample = ref.calibrate(3.0)
    times = []
    for s in range(args.warmup + args.steps):
        x, z = ref.patterns(s * sample, sample)
        sec, _ = ref.run(x, z)
        if s >= args.warmup:
            times.append(sec)
    total = float(np.sum(times))
    value = sample * args.steps / total
    desc = "%d frames per step of the same Philox stream (frame ids from 0), %d OpenMP threads, g++ -O2 -fopenmp" % (
        sample, ref.cores)
    line = {"impl": "reference", "metric": METRIC, "value": value, "unit": "frames/s", "n_gpus": args.gpus,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * total / args.steps,
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": workload_config(args.gpus, FRAMES_PER_STEP),
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
    F = args.frames

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    code = q.Code.qc(*CODE)
    dec = q.Decoder(code, local, F)
    stream = torch.cuda.Stream()
    dec.set_stream(stream.cuda_stream)
    info = [dec.launch_info(0), dec.launch_info(1)]

    def first_frame(step):  # global frame ids: disjoint Philox streams per (step, rank)
        return (step * world + rank) * F

    # ---- device-resident measurement ("value") ---------------------------------------------------------------
    for s in range(args.warmup):
        dec.get_statistics_depolarizing(SEED, first_frame(10_000 + s), F, P, MAXIT)
    dec.get_timing(reset=True)
    dec.enable_timing(True)
    sampler = ClockSampler(local)
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    counters = np.zeros(q.NUM_COUNTERS, np.uint64)
    barrier()
    sampler.start()
    ev0.record(stream)
    for s in range(args.steps):
        counters += dec.get_statistics_depolarizing(SEED, first_frame(s), F, P, MAXIT)["counters"]
    ev1.record(stream)
    barrier()
    clocks = sampler.result()
    ms_total = ev0.elapsed_time(ev1)
    kms, klaunch = dec.get_timing(reset=True)
    dec.enable_timing(False)

    # the only data collective: one NCCL all-reduce (sum) of the counter vector; the time is max-reduced over ranks
    from qec_ldpc_b200.sharding import allreduce_counters, allreduce_max
    ms_total = allreduce_max(ms_total, "cuda")
    gc = allreduce_counters(counters, "cuda")

    # ---- end to end through the C ABI with HOST buffers ("e2e") ----------------------------------------------
    # DecoderGPU::GetStats(..., xErrors, zErrors) (DecoderGPU.h:193): pre-generated patterns in host memory, in the
    # reference's layout (one int per qubit, frame-major); H2D copies, decode and the D2H counter read are timed.
    n = code.n
    code_nw = (n + 31) // 32
    xh = torch.empty((F, n), dtype=torch.int32, pin_memory=True)
    zh = torch.empty((F, n), dtype=torch.int32, pin_memory=True)
    sl = 100_000
    for off in range(0, F, sl):
        cnt = min(sl, F - off)
        x, z, _, _ = dec.debug_generate(SEED, first_frame(0) + off, cnt, P)
        xh[off:off + cnt] = torch.from_numpy(x)
        zh[off:off + cnt] = torch.from_numpy(z)
    # The library packs the rows to bits with host threads before the copy (1/32 of the bytes cross the link); with
    # many ranks per box there are too few cores per rank for that, and it copies the raw rows instead (threads = 0).
    local_world = int(os.environ.get("LOCAL_WORLD_SIZE", n_gpus))
    host_threads = min(16, len(os.sched_getaffinity(0)) // max(1, local_world))
    if host_threads < 6:
        host_threads = 0

    def time_i32(threads):
        dec.set_host_threads(threads)
        out = None
        for s in range(max(1, min(args.warmup, 2))):
            dec.get_stats_from_errors_ptr(xh.data_ptr(), zh.data_ptr(), F, P, MAXIT, elem=4)
        barrier()
        t0 = time.perf_counter()
        for s in range(args.steps):
            out = dec.get_stats_from_errors_ptr(xh.data_ptr(), zh.data_ptr(), F, P, MAXIT, elem=4)
        torch.cuda.synchronize()
        return allreduce_max(time.perf_counter() - t0, "cuda"), out

    e2e_raw_s, e2e_counters = time_i32(0)
    if host_threads > 0:
        e2e_s, packed_counters = time_i32(host_threads)
        assert np.array_equal(packed_counters, e2e_counters)
    else:
        e2e_s = e2e_raw_s
    # same patterns as step 0 of the device-resident run => same counters (checked on every rank)
    chk = dec.get_statistics_depolarizing(SEED, first_frame(0), F, P, MAXIT)["counters"]
    assert np.array_equal(chk, e2e_counters), "host-buffer path and device-generated path disagree"

    # two more end-to-end views, reported next to the primary one: (a) the same call with one BYTE per qubit
    # (qldpc_get_stats_from_errors_u8: a quarter of the PCIe traffic), (b) the Monte-Carlo call itself,
    # qldpc_get_statistics_depolarizing (what GetStatistics(W, COUNT, p, MAXIT) is in the reference's driver,
    # main.cu:101): scalars in, counters out, errors generated on the device -- timed on the host clock.
    xb, zb = xh.to(torch.uint8).pin_memory(), zh.to(torch.uint8).pin_memory()
    dec.get_stats_from_errors_ptr(xb.data_ptr(), zb.data_ptr(), F, P, MAXIT, elem=1)
    barrier()
    t0 = time.perf_counter()
    for s in range(args.steps):
        u8_counters = dec.get_stats_from_errors_ptr(xb.data_ptr(), zb.data_ptr(), F, P, MAXIT, elem=1)
    torch.cuda.synchronize()
    e2e_u8_s = allreduce_max(time.perf_counter() - t0, "cuda")
    assert np.array_equal(chk, u8_counters)
    barrier()
    t0 = time.perf_counter()
    for s in range(args.steps):
        dec.get_statistics_depolarizing(SEED, first_frame(s), F, P, MAXIT)
    torch.cuda.synchronize()
    e2e_gs_s = allreduce_max(time.perf_counter() - t0, "cuda")
    del xb, zb

    if rank == 0:
        frames = int(gc[0])
        eu = int(gc[9]) * code.EX + int(gc[10]) * code.EZ
        secs = ms_total * 1e-3
        value = frames / secs
        # roofline of the dominant kernel (bp_tile_kernel, X and Z launches), rank 0's launches
        k_eu = int(counters[9]) * code.EX + int(counters[10]) * code.EZ
        bp_ms = kms["bp_x"] + kms["bp_z"]
        bp_launches = klaunch["bp_x"] + klaunch["bp_z"]
        achieved = k_eu * BYTES_PER_EDGE_UPDATE / (bp_ms * 1e-3) / 1e9
        peaks_path = os.path.join(ROOT, "MEASURED_PEAKS.json")
        if os.path.exists(peaks_path):
            peak, peak_src = float(json.load(open(peaks_path))["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
        else:
            peak, peak_src = 6650.0, "fallback (B200_PROFILING.md)"
        traffic = None
        tpath = os.path.join(ROOT, "profiles", "traffic.json")
        if os.path.exists(tpath):
            traffic = json.load(open(tpath)).get("dram_bytes_per_launch")
        sm_mhz = clocks["sm_mhz"] or 1965.0
        smem_nominal = 148 * 128 * sm_mhz * 1e6 / 1e9  # GB/s at the SM clock sampled under load
        smem_peak, smem_src = smem_nominal, "nominal 148 SM x 128 B/clk x %.0f MHz (sampled under load)" % sm_mhz
        spath = os.path.join(ROOT, "profiles", "smem_peak.json")
        if os.path.exists(spath):  # measured by tools/micro/smem_bw.cu (LDS.128 + STS.128, the kernel's 1:1 mix)
            smem_peak = float(json.load(open(spath))["roofline_denominator_gbs"])
            smem_src = "measured (profiles/smem_peak.json, tools/micro/smem_bw.cu); nominal %.0f GB/s" % smem_nominal
        step_kernel_ms = sum(kms.values())

        def fp32_ops(dc, dv):  # FP32 instructions x lanes per edge-update of the reference's arithmetic
            check = dc * (dc - 1) / 2.0 + (dc - 2) + 2 * dc          # exclusive products with shared prefix, 1-2q, r
            var = dv * (dv - 1) + 2 * (dv - 1) + dv + dv + 5 * dv    # P and Q chains, prefixes, 1-p, Q+P, division
            return check / dc + var / dv
        fp32_total = (int(counters[9]) * code.EX * fp32_ops(code.dcX, code.dvX)
                      + int(counters[10]) * code.EZ * fp32_ops(code.dcZ, code.dvZ))
        fp32_achieved = fp32_total / (bp_ms * 1e-3) / 1e9
        fp32_peak = 148 * 128 * sm_mhz * 1e6 / 1e9
        line = {
            "metric": METRIC, "value": value, "unit": "frames/s", "n_gpus": n_gpus, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": ms_total / args.steps, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "f32", "data": "synthetic", "config": workload_config(n_gpus, F),
            "edge_updates_per_s": eu / secs,
            "mean_iterations": {"x": int(gc[9]) / frames, "z": int(gc[10]) / frames},
            "frame_error_rate": 1.0 - int(gc[3]) / frames,
            "counters": dict(zip(q.COUNTER_NAMES, [int(v) for v in gc])),
            "roofline": {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                         "traffic": traffic, "peak_source": peak_src, "kernel": "bp_tile_kernel (X + Z launches)",
                         "algorithmic_bytes_per_edge_update": BYTES_PER_EDGE_UPDATE,
                         "edge_updates_per_launch": k_eu / max(bp_launches, 1),
                         "avg_launch_ms": bp_ms / max(bp_launches, 1),
                         "kernel_share_of_step": bp_ms / max(step_kernel_ms, 1e-9),
                         "note": "messages stay in shared memory, so the HBM-denominated fraction may exceed 1; "
                                 "see smem_roofline for the tier that actually holds the state"},
            "smem_roofline": {"bound": "smem", "achieved": achieved, "peak": smem_peak, "unit": "GB/s",
                              "frac": achieved / smem_peak,
                              "peak_source": smem_src},
            "fp32_roofline": {"bound": "fp32 pipe", "achieved": fp32_achieved, "peak": fp32_peak,
                              "unit": "G lane-instructions/s (an FMA counts once)", "frac": fp32_achieved / fp32_peak,
                              "peak_source": "148 SM x 128 FP32 lanes x %.0f MHz (sampled under load)" % sm_mhz,
                              "note": "operation count fixed by the reference's arithmetic order (bit-exact parity)"},
            "kernel_ms_per_step": {k: v / args.steps for k, v in kms.items()},
            "launch": {"x": info[0], "z": info[1]},
            "e2e": {"value": F * n_gpus * args.steps / e2e_s, "unit": "frames/s",
                    "h2d_bytes_per_step": 2 * F * (code_nw * 4 if host_threads > 0 else n * 4),
                    "d2h_bytes_per_step": q.NUM_COUNTERS * 8,
                    "host_buffer_bytes_per_step": 2 * F * n * 4, "host_threads": host_threads,
                    "api": "qldpc_get_stats_from_errors_i32 (DecoderGPU::GetStats layout, pinned host int32)",
                    "note": "the reference's layout spends one int32 per qubit (4880 B per frame); the library's host "
                            "threads pack the rows to bits before the copy (host_threads > 0), otherwise the raw rows "
                            "cross the link (raw_rows_over_link: bound by the H2D link, about 51 GB/s per GPU); "
                            "u8_patterns and device_generated show the other two entry points",
                    "ms_per_step": 1e3 * e2e_s / args.steps,
                    "raw_rows_over_link": {"value": F * n_gpus * args.steps / e2e_raw_s, "unit": "frames/s",
                                           "h2d_bytes_per_step": 2 * F * n * 4, "host_threads": 0},
                    "u8_patterns": {"value": F * n_gpus * args.steps / e2e_u8_s, "unit": "frames/s",
                                    "h2d_bytes_per_step": 2 * F * (code_nw * 4 if host_threads > 0 else n),
                                    "host_buffer_bytes_per_step": 2 * F * n, "host_threads": host_threads,
                                    "api": "qldpc_get_stats_from_errors_u8"},
                    "device_generated": {"value": F * n_gpus * args.steps / e2e_gs_s, "unit": "frames/s",
                                         "h2d_bytes_per_step": 0, "d2h_bytes_per_step": q.NUM_COUNTERS * 8,
                                         "api": "qldpc_get_statistics_depolarizing (host wall clock)"}},
            "gpu_launches": int(sum(klaunch.values())),
            "clocks": clocks,
        }
        if n_gpus == 1 and not args.no_cpu:
            ref = CpuReference()
            sample = ref.calibrate(args.cpu_seconds)
            x, z = ref.patterns(0, sample)
            sec, corrected = ref.run(x, z)
            line["cpu_baseline"] = {
                "value": sample / sec, "unit": "frames/s", "cores": ref.cores, "kind": ref.kind,
                "sample": "first %d frames of the same Philox stream (seed %d), %.1f s wall, corrected %d"
                          % (sample, SEED, sec, corrected)}
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
    ap.add_argument("--frames", type=int, default=FRAMES_PER_STEP, help="frames per GPU per step")
    ap.add_argument("--no-cpu", action="store_true", help="skip the cpu_baseline leg")
    ap.add_argument("--cpu-seconds", type=float, default=12.0)
    args = ap.parse_args()
    if args.impl == "reference":
        run_reference(args)
    else:
        run_native(args)


if __name__ == "__main__":
    main()
