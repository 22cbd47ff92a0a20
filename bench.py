#!/usr/bin/env python
"""bench.py — rotated IoU Gpairs/s (headline) + NMS cands/s + FRM GB/s on B200, one JSON line.

    python bench.py [--gpus N] [--steps K] [--warmup W]                  (N > 1: launched by torchrun)
    python bench.py --impl reference [--gpus N] [--steps K] [--warmup W] (the reference's CPU implementation)

A "step" is one pass of the hot path over one batch: BASELINE.json configs[2], the rotated-IoU microbench —
RBboxOverlaps2D_v1 on 1,000 GT x 200,000 anchors per GPU (synthetic rotated boxes), through
the C ABI (r3g_iou_matrix_f32: one prepare kernel + the pair kernel).  The anchor axis is the shard axis: every
rank owns 200k anchors (weak scaling), GT replicated, no data-path collective.
  value     Gpairs/s, inputs resident in HBM, K steps timed with CUDA events between barriers, max over ranks
  e2e       same metric through the Python plugin API with HOST buffers: pinned H2D of both box sets + D2H of the
            (1000 x 200000) result inside the timed region; e2e.ceiling = the bare copies alone (platform limit)
  roofline  HBM store bound: 4 B per pair / mean duration of the pair kernel alone (r3g_iou_matrix_prepared_f32),
            against MEASURED_PEAKS.json hbm_gbs; traffic = ncu DRAM bytes of that kernel (profiles/r02_kernel_traffic.json)
  cpu_baseline  the reference's own host geometry (oracle/_ref, unmodified reference sources) on a bounded sample
Side sections.  EVERY rank runs `nms_batch` (configs[3]: its 8 images x K candidates x 15 classes, one launch sequence),
`train_step_hot_path` (configs[4]: 8 patches per GPU, two fused assignments + FRM forward / backward) and `assign_e2e`
(host boxes in, assignment out): times are the max over ranks, throughputs the whole-job aggregate.  N > 1 adds
`collectives` (the two harness exchanges, steady state).  Rank 0 adds the per-variant IoU, the single-image NMS sweep with its
roofline (reference pair count, measured pass fractions, both floors), FRM, the dense-head tail, the reference's own CUDA
kernels timed in the same run (`reference_cuda`) and the CPU baselines of NMS / FRM.
"""
import argparse
import ctypes as C
import json
import os
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

GT, ANCHORS = 1000, 200000
VARIANT = "v1"
METRIC, UNIT = "rotated_iou_pairs_per_s", "Gpairs/s"
WORKLOAD = ("configs[2] rotated IoU microbench: RBboxOverlaps2D_v1, 1000 GT x 200000 anchors per GPU "
            "(anchor axis sharded across ranks, GT replicated)")
AR = {'v1': (-np.pi / 2, 0), 'v2': (-np.pi / 4, 3 * np.pi / 4), 'v3': (-np.pi / 2, np.pi / 2)}
NMS_SIZES = (2000, 8000, 20000, 80000, 200000)
W_FULL, W_REJ = 224.0, 8.0          # FP32 flop of a clipped pair / of a pair rejected by the circumradius test


def rand_obb(n, seed, version='v1', lo=8, hi=512, span=1024):
    rng = np.random.default_rng(seed)
    cx = rng.uniform(0, span, n); cy = rng.uniform(0, span, n)
    w = np.exp(rng.uniform(np.log(lo), np.log(hi), n)); h = np.exp(rng.uniform(np.log(lo), np.log(hi), n))
    a = rng.uniform(*AR[version], n)
    return np.stack([cx, cy, w, h, a], 1).astype(np.float32)


def clustered(K, seed, version='v1', ncls=15):
    rng = np.random.default_rng(seed)
    seeds = rand_obb(max(K // 10, 1), seed + 1000, version, 12, 200)
    idx = rng.integers(0, len(seeds), K)
    b = seeds[idx].copy()
    b[:, 0:2] += rng.normal(0, 4, (K, 2)); b[:, 4] += rng.normal(0, 0.05, K)
    b[:, 2:4] *= np.exp(rng.normal(0, 0.1, (K, 2)))
    return b.astype(np.float32), rng.permutation(np.linspace(0.05, 1, K)).astype(np.float32), (idx % ncls).astype(np.int64)


def measured_peaks():
    """(HBM GB/s, FP32 TFLOP/s, source).  FP32 has no measured entry: 148 SMs x 128 lanes x 2 flop x the recorded max SM clock."""
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        d = json.load(open(p))
        mhz = float(d.get("sm_max_mhz", 1965.0))
        return float(d["hbm_gbs"]), 148 * 128 * 2 * mhz * 1e6 / 1e12, "measured (MEASURED_PEAKS.json)"
    return 6650.0, 148 * 128 * 2 * 1.965e9 / 1e12, "fallback (B200_PROFILING.md)"


def kernel_traffic(name):
    """ncu DRAM bytes per launch of a kernel on its bench workload (scripts/ncu_traffic.py -> profiles/r02_kernel_traffic.json)."""
    p = os.path.join(ROOT, "profiles", "r02_kernel_traffic.json")
    try:
        rec = json.load(open(p))[name]
        return int(rec["dram_bytes"]), rec.get("source", "profiles/r02_kernel_traffic.json")
    except Exception:  # noqa: BLE001
        return None, "no ncu capture committed for this kernel"


# ------------------------------------------------------------------------------------------ clocks
class ClockSampler:
    """Samples SM clock + throttle reasons during the timed region (pynvml, 20 ms period)."""

    def __init__(self, index):
        self.samples, self.reasons, self.max_mhz, self._stop, self.ok = [], set(), None, threading.Event(), False
        try:
            import pynvml
            pynvml.nvmlInit()
            self.nv = pynvml
            self.h = pynvml.nvmlDeviceGetHandleByIndex(index)
            self.max_mhz = int(pynvml.nvmlDeviceGetMaxClockInfo(self.h, pynvml.NVML_CLOCK_SM))
            self.ok = True
        except Exception:  # noqa: BLE001
            self.ok = False
        self.t = threading.Thread(target=self._run, daemon=True)

    def _run(self):
        nv = self.nv
        names = {"hw_slowdown": 0x8, "sw_power_cap": 0x4, "hw_thermal_slowdown": 0x40, "sw_thermal_slowdown": 0x20,
                 "hw_power_brake": 0x80, "sync_boost": 0x10, "applications_clocks_setting": 0x2}
        while not self._stop.is_set():
            try:
                self.samples.append(int(nv.nvmlDeviceGetClockInfo(self.h, nv.NVML_CLOCK_SM)))
                try:
                    r = int(nv.nvmlDeviceGetCurrentClocksEventReasons(self.h))
                except Exception:  # noqa: BLE001
                    r = int(nv.nvmlDeviceGetCurrentClocksThrottleReasons(self.h))
                for k, bit in names.items():
                    if r & bit:
                        self.reasons.add(k)
            except Exception:  # noqa: BLE001
                pass
            time.sleep(0.02)

    def __enter__(self):
        if self.ok:
            self.t.start()
        return self

    def __exit__(self, *a):
        self._stop.set()
        if self.ok:
            self.t.join(1.0)

    def summary(self):
        if not self.samples:
            return {"sm_mhz": None, "sm_max_mhz": self.max_mhz, "reasons": ["unavailable"]}
        return {"sm_mhz": int(np.median(self.samples)), "sm_max_mhz": self.max_mhz, "reasons": sorted(self.reasons),
                "samples": len(self.samples)}


# ------------------------------------------------------------------------------------------ reference CPU arm
def ref_cpu_fn():
    """(callable(b1, b2) -> IoU matrix, kind): the reference's own host geometry when oracle/_ref travelled here,
    else the oracle port."""
    try:
        from oracle import ref
        if ref.available("libref_v1.so"):
            ref.v1_iou(rand_obb(2, 0), rand_obb(2, 1))
            return (lambda a, b: ref.v1_iou(a, b)), "reference"
    except Exception:  # noqa: BLE001
        pass
    from oracle import port
    return (lambda a, b: port.iou_matrix(a, b, VARIANT)), "port"


def cpu_pairs_per_s(anchors_per_thread, threads, seed=100):
    """All `threads` host threads, each on its own anchor shard x the 1,000 GT (ctypes releases the GIL)."""
    from concurrent.futures import ThreadPoolExecutor
    fn, kind = ref_cpu_fn()
    gt = rand_obb(GT, 1, VARIANT)
    shards = [rand_obb(anchors_per_thread, seed + i, VARIANT) for i in range(threads)]
    t0 = time.perf_counter()
    with ThreadPoolExecutor(threads) as ex:
        list(ex.map(lambda a: fn(gt, a), shards))
    dt = time.perf_counter() - t0
    return GT * anchors_per_thread * threads / dt, dt, kind


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    cores = os.cpu_count() or 1
    # calibrate, then size the per-step sample so that (steps + warmup) steps finish in ~2 minutes
    rate, _, kind = cpu_pairs_per_s(500, cores)
    budget_s = 110.0 / max(args.steps + args.warmup, 1)
    apt = int(max(20, min(40000, rate * budget_s / (GT * cores))))
    for _ in range(args.warmup):
        cpu_pairs_per_s(apt, cores)
    t0 = time.perf_counter()
    for i in range(args.steps):
        cpu_pairs_per_s(apt, cores, seed=200 + i)
    dt = time.perf_counter() - t0
    pairs = GT * apt * cores * args.steps
    val = pairs / dt / 1e9
    sample = f"{GT} GT x {apt * cores} anchors per step ({apt} per thread), {args.steps} steps"
    _emit({
        "impl": "reference", "metric": METRIC, "value": val, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": dt / max(args.steps, 1) * 1e3, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": WORKLOAD, "variant": VARIANT, "gt": GT, "anchors_per_gpu": ANCHORS,
                   "strict_reference_parity": True,
                   "sample": "each step is a bounded CPU sample of the same boxes: " + sample},
        "cpu_baseline": {"value": val, "unit": UNIT, "cores": cores, "kind": kind, "sample": sample},
        "e2e": {"value": val, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    })


def pin_to_gpu_numa_node(index):
    """Run this rank (and first-touch its pinned host buffers) on the CPUs next to its GPU: the end-to-end number moves
    800 MB per step over PCIe, and with several ranks on one host the cross-socket path is the slow one."""
    try:
        import pynvml
        pynvml.nvmlInit()
        h = pynvml.nvmlDeviceGetHandleByIndex(index)
        words = (os.cpu_count() + 63) // 64
        mask = pynvml.nvmlDeviceGetCpuAffinity(h, words)
        cpus = {64 * w + b for w, m in enumerate(mask) for b in range(64) if (int(m) >> b) & 1}
        cpus &= set(os.sched_getaffinity(0))
        if cpus:
            os.sched_setaffinity(0, cpus)
            return len(cpus)
    except Exception:  # noqa: BLE001
        pass
    return None


# ------------------------------------------------------------------------------------------ GPU arm
class Ranks:
    """Barrier / max-over-ranks helpers (torch.distributed over NCCL when world > 1)."""

    def __init__(self, torch, dist, dev, world, rank):
        self.torch, self.dist, self.dev, self.world, self.rank = torch, dist, dev, world, rank

    def barrier(self):
        if self.world > 1:
            self.dist.barrier()
        self.torch.cuda.synchronize()

    def max(self, x):
        if self.world > 1:
            t = self.torch.tensor([x], dtype=self.torch.float64, device=self.dev)
            self.dist.all_reduce(t, op=self.dist.ReduceOp.MAX)
            return float(t.item())
        return float(x)

    def time(self, fn, iters, warm=3):
        """ms per call: CUDA events around `iters` calls between barriers, max over ranks."""
        torch = self.torch
        for _ in range(warm):
            fn()
        self.barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(iters):
            fn()
        e1.record()
        self.barrier()
        return self.max(e0.elapsed_time(e1)) / iters


def run_gpu(args):
    import torch
    import torch.distributed as dist
    import r3det_b200 as R
    from r3det_b200 import _lib as L

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    numa = pin_to_gpu_numa_node(local)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=dev)
    rk = Ranks(torch, dist, dev, world, rank)

    lib = L.lib()
    gt_h = torch.from_numpy(rand_obb(GT, 1, VARIANT)).pin_memory()
    an_h = torch.from_numpy(rand_obb(ANCHORS, 1000 + rank, VARIANT)).pin_memory()       # this rank's anchor shard
    gt, an = gt_h.to(dev), an_h.to(dev)
    out = torch.empty((GT, ANCHORS), dtype=torch.float32, device=dev)
    nbytes = C.c_size_t(0)
    L.check(lib.r3g_iou_workspace_bytes(GT, ANCHORS, C.byref(nbytes)))
    ws = torch.empty(nbytes.value, dtype=torch.uint8, device=dev)
    stream = torch.cuda.current_stream(dev)
    sp = C.c_void_p(stream.cuda_stream)
    flags = L.FLAG_STRICT
    pairs = GT * ANCHORS

    def step():
        L.check(lib.r3g_iou_matrix_f32(L.ptr(gt), GT, 5, L.ptr(an), ANCHORS, 5, L.V[VARIANT], 0, flags,
                                       L.ptr(out), L.ptr(ws), ws.numel(), sp))

    def pair_kernel_only():
        L.check(lib.r3g_iou_matrix_prepared_f32(L.ptr(gt), GT, 5, L.ptr(an), ANCHORS, 5, L.V[VARIANT], 0, flags,
                                                L.ptr(out), L.ptr(ws), ws.numel(), sp))

    with ClockSampler(local) as clocks:
        for _ in range(max(args.warmup, 3)):
            step()
        rk.barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(stream)
        for _ in range(args.steps):
            step()
        e1.record(stream)
        rk.barrier()
        ms_total = rk.max(e0.elapsed_time(e1))
        # the dominant kernel alone (prepared boxes already in the workspace), same K launches
        k0, k1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        k0.record(stream)
        for _ in range(args.steps):
            pair_kernel_only()
        k1.record(stream)
        rk.barrier()
        ms_kernel = k0.elapsed_time(k1) / args.steps
    stats = ws[:32].view(torch.int64).cpu().numpy().tolist()

    # ---- e2e: host buffers through the plugin API (H2D of both box sets, D2H of the result, every step)
    res_h = torch.empty((GT, ANCHORS), dtype=torch.float32).pin_memory()
    calc = R.RBboxOverlaps2D_v1()
    e2e_steps = max(3, min(args.steps, 20))

    def e2e_step():
        g = gt_h.to(dev, non_blocking=True)
        a = an_h.to(dev, non_blocking=True)
        o = calc(g, a)
        res_h.copy_(o, non_blocking=True)

    def bare_copies():                      # the same bytes over PCIe without any kernel: the platform's ceiling for `e2e`
        gt_h.to(dev, non_blocking=True)
        an_h.to(dev, non_blocking=True)
        res_h.copy_(out, non_blocking=True)

    ms_e2e = rk.time(e2e_step, e2e_steps, warm=2)
    assert float(res_h[0].max()) >= 0.0
    ms_copy = rk.time(bare_copies, e2e_steps, warm=2)

    hbm, f32_peak, peak_src = measured_peaks()
    value = world * pairs * args.steps / (ms_total * 1e-3) / 1e9
    achieved = 4.0 * pairs / (ms_kernel * 1e-3) / 1e9
    traffic, traffic_src = kernel_traffic("iou_matrix_kernel")
    d2h = int(res_h.numel() * 4)
    line = {
        "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": max(args.warmup, 3),
        "ms_per_step": ms_total / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "f32", "data": "synthetic",
        "config": {"workload": WORKLOAD, "variant": VARIANT, "gt": GT, "anchors_per_gpu": ANCHORS, "strict_reference_parity": True,
                   "l2": "each step writes an 800 MB result (> 126 MB L2); no flush needed"},
        "e2e": {"value": world * pairs / (ms_e2e * 1e-3) / 1e9, "unit": UNIT,
                "h2d_bytes_per_step": int(gt_h.numel() * 4 + an_h.numel() * 4), "d2h_bytes_per_step": d2h,
                "ms_per_step": ms_e2e, "steps": e2e_steps,
                "ceiling": {"value": world * pairs / (ms_copy * 1e-3) / 1e9, "unit": UNIT, "ms_per_step": ms_copy,
                            "d2h_gbs_all_ranks": world * d2h / ms_copy / 1e6,
                            "what": "the same pinned H2D + D2H copies with no kernel in between, all ranks at once: the "
                                    "PCIe / host-memory limit of returning the dense matrix"}},
        "gpu_launches": 2 * args.steps,          # per step: prep_pair_kernel + iou_matrix_kernel
        "roofline": {"bound": "hbm", "achieved": achieved, "peak": hbm, "unit": "GB/s", "frac": achieved / hbm,
                     "traffic": traffic, "traffic_source": traffic_src, "peak_source": peak_src, "kernel": "iou_matrix_kernel<true, 0, false>",
                     "kernel_ms": ms_kernel, "algorithmic_bytes_per_launch": 4 * pairs,
                     "pairs_circle_pass": stats[0], "pairs_sat_pass": stats[1], "pairs_strict": stats[2],
                     "f_circle_pass": stats[0] / pairs,
                     "fp32_floor_ms": ((stats[0] / pairs) * W_FULL + (1 - stats[0] / pairs) * W_REJ) * pairs / (f32_peak * 1e12) * 1e3,
                     "fp32_peak_tflops": f32_peak, "fp32_peak_source": "148 SMs x 128 lanes x 2 flop x max SM clock (no measured FP32 entry)"},
        "clocks": clocks.summary(),
        "host_cpus_near_gpu": numa,
    }

    # ---- every rank: configs[3], configs[4] and the assigner end to end (max over ranks, whole-job aggregates)
    line["nms_batch"] = bench_nms_batch(torch, R, dev, rk)
    line["train_step_hot_path"] = bench_train_step(torch, R, dev, rk, dist)
    line["assign_e2e"] = bench_assign_e2e(torch, R, dev, rk, gt_h, an_h)
    if world > 1:
        line["collectives"] = bench_collectives(torch, dist, R, dev, rk)
    if rank == 0:
        line["iou_variants"] = bench_iou_variants(torch, R, dev)
        line["nms"] = bench_nms(torch, R, dev, hbm, f32_peak)
        line["frm"] = bench_frm(torch, R, dev, hbm)
        line["dense_tail"] = bench_dense_tail(torch, R, dev)
        line["reference_cuda"] = bench_reference_cuda(torch, dev, line)
        cores = os.cpu_count() or 1
        apt = 12000
        rate, dt, kind = cpu_pairs_per_s(apt, cores)
        line["cpu_baseline"] = {"value": rate / 1e9, "unit": UNIT, "cores": cores, "kind": kind,
                                "sample": f"{GT} GT x {apt * cores} anchors ({apt} per thread, {dt:.1f} s wall)"}
        line["nms"]["cpu_baseline"] = cpu_nms_baseline(cores)
        line["frm"]["cpu_baseline"] = cpu_frm_baseline()
    if world > 1:
        dist.barrier()
    if rank == 0:
        _emit(line)
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


def _time(torch, fn, iters, warm=3):
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(iters):
        fn()
    e1.record()
    e1.synchronize()
    return e0.elapsed_time(e1) / iters


def _graph_time(torch, fn, iters=20):
    """The same call sequence replayed from a CUDA graph (the library never syncs or allocates): device time without
    the Python wrapper's host overhead."""
    try:
        g = torch.cuda.CUDAGraph()
        with torch.cuda.graph(g):
            fn()
        return _time(torch, g.replay, iters)
    except Exception as e:  # noqa: BLE001
        return f"capture failed: {e}"


def bench_iou_variants(torch, R, dev):
    """configs[2] names all three angle conventions: the same 1000 x 200000 matrix for v1 / v2 / v3 (device-resident inputs,
    strict reference parity, whole op = one prepare launch + the pair kernel) and the IoF mode of v1."""
    out = {}
    for v, mode in (("v1", "iou"), ("v2", "iou"), ("v3", "iou"), ("v1", "iof")):
        gt = torch.from_numpy(rand_obb(GT, 1, v)).to(dev)
        an = torch.from_numpy(rand_obb(ANCHORS, 1000, v)).to(dev)
        ms = _time(torch, lambda: R.pairwise_iou(gt, an, v, mode), 30)
        out[f"{v}_{mode}"] = {"ms": ms, "gpairs_per_s": GT * ANCHORS / ms / 1e6}
    return out


def bench_assign_e2e(torch, R, dev, rk, gt_h, an_h):
    """SURVEY §8f rank 1, every rank on its anchor shard: the same 1000 x 200000 pairs consumed by the fused MaxIoUAssigner
    (pos 0.5 / neg 0.4 / min_pos 0, gt_max_assign_all) — no (G, A) matrix is stored.  `e2e` = pinned host boxes in, assignment
    (int64 per anchor) + max overlaps back on the host: the path's real end-to-end when the consumer is the assigner."""
    gt, an = gt_h.to(dev), an_h.to(dev)
    fn = lambda: R.max_iou_assign(gt, an, 0.5, 0.4, 0.0, True, True, VARIANT)
    ms = rk.time(fn, 50)
    res_i = torch.empty((ANCHORS,), dtype=torch.int64).pin_memory()
    res_f = torch.empty((ANCHORS,), dtype=torch.float32).pin_memory()

    def e2e():
        o = R.max_iou_assign(gt_h.to(dev, non_blocking=True), an_h.to(dev, non_blocking=True), 0.5, 0.4, 0.0, True, True, VARIANT)
        res_i.copy_(o.gt_inds, non_blocking=True); res_f.copy_(o.max_overlaps, non_blocking=True)

    ms_e2e = rk.time(e2e, 20)
    pairs = GT * ANCHORS * rk.world
    return {"ms": ms, "gpairs_per_s": pairs / ms / 1e6, "e2e_ms": ms_e2e, "e2e_gpairs_per_s": pairs / ms_e2e / 1e6,
            "d2h_bytes_per_step": int(res_i.numel() * 8 + res_f.numel() * 4), "pair_sweeps": 1, "ranks": rk.world,
            "num_pos_rank0": int((res_i > 0).sum())}


def _nms_roofline(K, labels, counters, ms_device, kept, hbm, f32_peak, sort_passes):
    """SURVEY §8d NMS model.  Pi = sum_c K_c (K_c - 1) / 2 is what a per-class triangular mask evaluates (the reference's own
    kernel does the full K^2 / 2); the rounds kernel evaluates only pairs of kept rows, so `pairs_tested` << Pi and the
    fraction against the reference floor may pass 1."""
    kc = np.bincount(labels).astype(np.float64)
    pi = float((kc * (kc - 1) / 2).sum())
    tested, circ, area, emu = (float(counters[i]) for i in range(4))
    f = circ / max(tested, 1.0)
    fp32_ref = (f * W_FULL + (1 - f) * W_REJ) * pi / (f32_peak * 1e12) * 1e3
    fp32_eval = (area * W_FULL + max(tested - area, 0.0) * W_REJ) / (f32_peak * 1e12) * 1e3
    mask_bytes = float(counters[5]) * 64 * 2 * 8 * 2                      # chunk-local mask words, written and read once
    bytes_ = 32.0 * K + 16.0 * K * sort_passes + mask_bytes + 8.0 * kept
    hbm_ms = bytes_ / (hbm * 1e9) * 1e3
    floor = max(fp32_ref, hbm_ms)
    return {"bound": "fp32" if fp32_ref >= hbm_ms else "hbm", "pairs_reference": pi, "pairs_tested": tested,
            "pairs_circle_pass": circ, "pairs_area_evaluated": area, "pairs_restated": emu, "f_circle_pass": f,
            "rounds": int(counters[4]), "fp32_floor_reference_ms": fp32_ref, "fp32_floor_evaluated_ms": fp32_eval,
            "hbm_floor_ms": hbm_ms, "algorithmic_bytes": bytes_, "device_ms": ms_device, "frac": floor / ms_device,
            "frac_of": "max(FP32 floor of the reference pair count, HBM floor) / device time (CUDA-graph replay)"}


def bench_nms(torch, R, dev, hbm, f32_peak):
    """configs[3], one image: clustered candidates x 15 DOTA classes, nms v1 (batched_rnms semantics), thr 0.1 — eager
    (with the Python wrapper), CUDA-graph replay (device time), work counters and the roofline per K."""
    from r3det_b200._nms_core import nms_device
    out = {"unit": "Mcands/s", "variant": "v1", "classes": 15, "iou_thr": 0.1, "sweep": {},
           "l2": "candidates fit in L2 (a latency / FP32-bound op: the HBM floor is the small one); nothing to flush"}
    for K in NMS_SIZES:
        b, s, l = clustered(K, 2, "v1")
        B, S, Lb = (torch.from_numpy(x).to(dev) for x in (b, s, l))
        scale = torch.tensor(float(b.max() + 1), device=dev)
        fn = lambda: nms_device(B, S, 0.1, "v1", labels=Lb, class_offset=scale, order_index=True, label_bits=4)
        st = {}
        keep, num = nms_device(B, S, 0.1, "v1", labels=Lb, class_offset=scale, order_index=True, label_bits=4, stats=st)
        torch.cuda.synchronize()
        counters = st["counters"].cpu().numpy()
        kept = int(num)
        ms = _time(torch, fn, 5 if K >= 80000 else 20)
        gms = _graph_time(torch, fn)
        rec = {"ms": ms, "mcands_per_s": K / ms / 1e3, "kept": kept, "graph_ms": gms}
        dev_ms = gms if isinstance(gms, float) else ms
        if isinstance(gms, float):
            rec["graph_mcands_per_s"] = K / gms / 1e3
        rec["roofline"] = _nms_roofline(K, l, counters, dev_ms, kept, hbm, f32_peak, 0 if K <= 16384 else 5)
        if K == 200000:
            t, src = kernel_traffic("nms_rounds_kernel_K200000")
            rec["roofline"]["traffic"], rec["roofline"]["traffic_source"] = t, src
        out["sweep"][str(K)] = rec
    return out


def bench_nms_batch(torch, R, dev, rk):
    """configs[3] as sharded: 64 images over 8 GPUs = 8 images per rank, K candidates x 15 classes each, ONE launch sequence per
    rank ((image, class) pairs are the segments).  Every rank runs its own 8 images; ms = max over ranks, Mcands/s = all ranks."""
    from r3det_b200._nms_core import nms_device
    out = {"unit": "Mcands/s", "images_per_rank": 8, "ranks": rk.world, "sweep": {}}
    for K in NMS_SIZES:
        imgs = [clustered(K, 100 + 8 * rk.rank + i, "v1") for i in range(8)]
        B = torch.from_numpy(np.concatenate([x[0] for x in imgs])).to(dev)
        S = torch.from_numpy(np.concatenate([x[1] for x in imgs])).to(dev)
        Lb = torch.from_numpy(np.concatenate([x[2] for x in imgs])).to(dev)
        bid = torch.arange(8, device=dev).repeat_interleave(K)
        scales = torch.tensor([float(x[0].max() + 1) for x in imgs], device=dev)
        fn = lambda: nms_device(B, S, 0.1, "v1", labels=Lb, class_offset=scales, order_index=True, batch_ids=bid, n_batches=8,
                                label_bits=4)
        keep, num = fn()
        ms = rk.time(fn, 5 if K >= 80000 else 10)
        rec = {"ms": ms, "mcands_per_s": rk.world * 8 * K / ms / 1e3, "kept_rank0": int(num.sum())}
        gms = _graph_time(torch, fn)                          # the same launch sequence replayed from a CUDA graph (rank-local)
        if isinstance(gms, float):
            rec["graph_ms_rank"] = gms
        out["sweep"][str(K)] = rec
        del B, S, Lb, bid
    return out


def _frm_inputs(torch, dev, seed, batch=8, channels=256):
    rng = np.random.default_rng(seed)
    xs, bts, scales = [], [], []
    for H, stride in ((128, 8), (64, 16), (32, 32), (16, 64), (8, 128)):
        xs.append(torch.randn((batch, channels, H, H), device=dev))
        ys_, xs_ = np.meshgrid(np.arange(H) * stride, np.arange(H) * stride, indexing="ij")
        ctr = np.stack([xs_, ys_], -1).reshape(-1, 2).astype(np.float32)
        bx = np.zeros((batch, H * H, 5), np.float32)
        bx[:, :, :2] = ctr[None] + rng.normal(0, stride, (batch, H * H, 2))
        bx[:, :, 2:4] = np.exp(rng.uniform(np.log(stride), np.log(8 * stride), (batch, H * H, 2)))
        bx[:, :, 4] = rng.uniform(-np.pi / 2, 0, (batch, H * H))
        bts.append(torch.from_numpy(bx.reshape(-1, 5)).to(dev)); scales.append(1.0 / stride)
    return xs, bts, scales


def bench_train_step(torch, R, dev, rk, dist):
    """configs[4] / configs[1]: the hot-path part of one R3Det training step, EVERY rank on its own 8 patches — per patch the
    assignment of 128 GT against the 196,416 RRetinaNet anchors and against the 21,824 refine-stage boxes (fused
    MaxIoUAssigner, v1, the 8 patches in one launch sequence), plus FRM forward + backward over the five FPN levels.
    Eager launches and the same sequence replayed from a CUDA graph; ms = max over ranks.  `ddp_grad_allreduce_ms` is what
    runs BESIDE it in the reference's distributed training (tools/dist_train.sh:7-9, torch DDP): the NCCL all-reduce of a
    37 M-parameter FP32 gradient (R3Det R50-FPN size) in 25 MB buckets — torch's, untouched, reported for context."""
    from r3det_b200.fr import frm_backward_multi, frm_forward_multi
    gts = [torch.from_numpy(rand_obb(128, 300 + 8 * rk.rank + i, "v1", 10, 300)).to(dev) for i in range(8)]
    anc = torch.from_numpy(rand_obb(196416, 400, "v1", 16, 512)).to(dev)
    refs = [torch.from_numpy(rand_obb(21824, 500 + 8 * rk.rank + i, "v1", 10, 400)).to(dev) for i in range(8)]
    xs, bts, scales = _frm_inputs(torch, dev, 12 + rk.rank)
    refs_b = torch.stack(refs)

    def step_per_image():
        for i in range(8):
            R.max_iou_assign(gts[i], anc, 0.5, 0.4, 0.0, True, True, "v1")
            R.max_iou_assign(gts[i], refs[i], 0.5, 0.4, 0.0, True, True, "v1")
        frm_forward_multi(xs, bts, scales, 1)
        frm_backward_multi(xs, bts, scales, 1)

    def step():
        R.max_iou_assign_batched(gts, anc, 0.5, 0.4, 0.0, True, True, "v1")          # the 8 patches in one launch sequence
        R.max_iou_assign_batched(gts, refs_b, 0.5, 0.4, 0.0, True, True, "v1")
        frm_forward_multi(xs, bts, scales, 1)
        frm_backward_multi(xs, bts, scales, 1)

    ms = rk.time(step, 10)
    pairs = 8 * 128 * (196416 + 21824)
    out = {"images_per_rank": 8, "ranks": rk.world, "gt_per_image": 128, "ms": ms, "pairs_per_rank": pairs,
           "images_per_s": rk.world * 8 / ms * 1e3,
           "per_image_assign_calls_ms": rk.time(step_per_image, 5)}
    try:
        g = torch.cuda.CUDAGraph()
        with torch.cuda.graph(g):
            step()
        out["graph_ms"] = rk.time(g.replay, 10)
    except Exception as e:  # noqa: BLE001
        out["graph_ms"] = f"capture failed: {e}"
        rk.barrier()
    if rk.world > 1:
        buckets = [torch.randn((25 * 1024 * 1024 // 4,), device=dev) for _ in range(6)]      # 6 x 25 MB ~ 37 M parameters

        def allreduce():
            for b in buckets:
                dist.all_reduce(b)

        out["ddp_grad_allreduce_ms"] = rk.time(allreduce, 10)
        out["ddp_grad_bytes"] = int(sum(b.numel() * 4 for b in buckets))
    return out


def bench_collectives(torch, dist, R, dev, rk):
    """N > 1 only: the two harness-level exchanges of the path over NCCL, in STEADY STATE (warm communicator, 20
    iterations, max over ranks):
      assigner_stats   — every rank assigns its anchor block with the FUSED assigner (no overlap matrix), then ONE all_gather of
                         the G packed per-GT maxima + pos / neg counters (sharding.assigner_stats); timed with and without the
                         assignment itself;
      keep lists       — every rank runs batched NMS on its 8 images and publishes padded (2000 x 7) records with ONE
                         all_gather (sharding.gather_padded_records); `pipelined` issues the gather asynchronously (NCCL's own
                         stream) under the NMS of the next batch.
    Results are checked for consistency across ranks and against the unsharded answer."""
    from r3det_b200 import sharding
    from r3det_b200._nms_core import nms_device
    world, rank = rk.world, rk.rank
    gt = torch.from_numpy(rand_obb(GT, 1, VARIANT)).to(dev)
    anchors = torch.from_numpy(rand_obb(ANCHORS, 7, VARIANT)).to(dev)          # same anchors everywhere; each rank takes its rows
    assign_fn = lambda g, a: R.max_iou_assign(g, a, 0.5, 0.4, 0.0, True, True, VARIANT)

    def stats_full():
        o, lo, hi = sharding.sharded_assign(gt, anchors, assign_fn)
        return sharding.assigner_stats(o.gt_max_overlaps, o.gt_argmax_overlaps, o.max_overlaps, lo, 0.5, 0.4)

    o, lo, hi = sharding.sharded_assign(gt, anchors, assign_fn)
    stats_only = lambda: sharding.assigner_stats(o.gt_max_overlaps, o.gt_argmax_overlaps, o.max_overlaps, lo, 0.5, 0.4, sync=False)
    gmax, garg, npos, nneg = stats_full()
    ms_stats_full = rk.time(stats_full, 20)
    ms_stats_only = rk.time(stats_only, 20)

    # image-sharded NMS: 8 images per rank, padded records built on the device
    K, M = 2000, 2000
    images = 8 * world
    imgs = [clustered(K, 50 + 8 * rank + i, VARIANT) for i in range(8)]
    B = torch.from_numpy(np.concatenate([x[0] for x in imgs])).to(dev)
    S = torch.from_numpy(np.concatenate([x[1] for x in imgs])).to(dev)
    Lb = torch.from_numpy(np.concatenate([x[2] for x in imgs])).to(dev)
    bid = torch.arange(8, device=dev).repeat_interleave(K)
    scales = torch.tensor([float(x[0].max() + 1) for x in imgs], device=dev)

    def nms_records():
        keep, num = nms_device(B, S, 0.1, VARIANT, labels=Lb, class_offset=scales, order_index=True, batch_ids=bid, n_batches=8,
                               label_bits=4)
        return R.pack_keep_records(B, S, Lb, keep, num, bid, 8, M)

    dets, labels, counts = nms_records()
    all_dets, all_labels = sharding.gather_padded_records(dets, labels, counts, images)       # per-image lists (one host read)
    gather = lambda: sharding.gather_padded_records(dets, labels, counts, images, padded=True)  # padded device tensors, no host read
    ms_gather = rk.time(gather, 20)
    ms_gather_lists = rk.time(lambda: sharding.gather_padded_records(dets, labels, counts, images), 20)

    def nms_then_gather():
        d, l, c = nms_records()
        sharding.gather_padded_records(d, l, c, images, padded=True)

    pending = []

    def pipelined():                          # gather of batch i in flight under the NMS of batch i + 1
        d, l, c = nms_records()
        h = sharding.gather_padded_records(d, l, c, images, async_op=True)
        if pending:
            pending.pop().wait_padded()
        pending.append(h)

    ms_serial = rk.time(nms_then_gather, 20)
    ms_pipe = rk.time(pipelined, 20)
    while pending:
        pending.pop().wait_padded()
    ms_nms = rk.time(nms_records, 20)

    # every rank must hold the same global view
    chk = torch.tensor([float(gmax.double().sum()), float(garg.double().sum()), float(npos), float(nneg),
                        float(sum(d.double().sum() for d in all_dets))], dtype=torch.float64, device=dev)
    ref = chk.clone(); dist.broadcast(ref, 0)
    ok = bool(torch.equal(chk, ref)) and len(all_dets) == images
    if rank == 0:
        full = R.pairwise_iou(gt, anchors, VARIANT)
        ok = ok and bool(torch.equal(gmax, full.max(dim=1)[0])) and bool(torch.equal(garg, full.max(dim=1)[1]))
        amax = full.max(dim=0)[0]
        ok = ok and npos == int((amax >= 0.5).sum()) and nneg == int(((amax >= 0) & (amax < 0.4)).sum())
    return {"consistent": ok, "images": images, "num_pos": npos, "num_neg": nneg, "iterations": 20,
            "assigner_stats_ms": ms_stats_only, "sharded_assign_plus_stats_ms": ms_stats_full,
            "gather_keep_lists_ms": ms_gather, "gather_keep_lists_as_python_lists_ms": ms_gather_lists, "nms_batch8_ms": ms_nms, "nms_then_gather_ms": ms_serial,
            "nms_gather_pipelined_ms": ms_pipe, "payload_bytes_per_rank": int(8 * (M * 7 + 1) * 4),
            "note": "steady state: warm NCCL communicator, CUDA events, max over ranks"}


def bench_dense_tail(torch, R, dev):
    """SURVEY §8f ranks 2-3 on configs[1] shapes: batch 8, 1024^2 patch, 5 FPN levels, 15 classes.
    get_bboxes: 9 anchors/location (196,416 rows/image), nms_pre 2000, score_thr 0.05, nms v1 thr 0.1, max 2000/img
    (select + decode + multiclass NMS, the whole batch in one launch sequence).  filter / refine: the FRM prologue."""
    rng = np.random.default_rng(9)
    Bn, A, Cn = 8, 9, 15
    cls, reg, anc, reg1 = [], [], [], []
    for H, stride in ((128, 8), (64, 16), (32, 32), (16, 64), (8, 128)):
        # focal-loss style logits: background prior 0.01 with a sparse set of confident locations
        c = rng.normal(-4.6, 1.0, (Bn, A * Cn, H, H)).astype(np.float32)
        hot = rng.random((Bn, A * Cn, H, H)) < 2e-4
        c[hot] = rng.normal(1.0, 1.0, int(hot.sum())).astype(np.float32)
        cls.append(torch.from_numpy(c).to(dev))
        reg.append(torch.from_numpy(rng.normal(0, 0.2, (Bn, A * 5, H, H)).astype(np.float32)).to(dev))
        reg1.append(torch.from_numpy(rng.normal(0, 0.2, (Bn, 5, H, H)).astype(np.float32)).to(dev))
        ys, xs = np.meshgrid(np.arange(H), np.arange(H), indexing="ij")
        ctr = (np.stack([xs, ys], -1).reshape(-1, 1, 2) * stride + stride / 2).astype(np.float32)
        wh = (stride * 4 * np.array([[1, 1], [1.4, 0.7], [0.7, 1.4]], np.float32)[None].repeat(3, 1)
              * np.array([1, 1, 1, 1.26, 1.26, 1.26, 1.59, 1.59, 1.59], np.float32)[None, :, None])
        a = np.concatenate([np.broadcast_to(ctr, (H * H, A, 2)), np.broadcast_to(wh, (H * H, A, 2)), np.zeros((H * H, A, 1), np.float32)], -1)
        anc.append(torch.from_numpy(np.ascontiguousarray(a.reshape(-1, 5))).to(dev))
    coder = R.DeltaXYWHAOBBoxCoder((0.,) * 5, (1.,) * 5, angle_range="v1")
    metas = [dict(img_shape=(1024, 1024, 3), scale_factor=np.ones(4, np.float32))] * Bn
    cfg = dict(nms_pre=2000, min_bbox_size=0, score_thr=0.05, nms=dict(type="v1", iou_thr=0.1), max_per_img=2000)
    sel = lambda: R.select_decode(cls, reg, anc, coder, 2000, [m["img_shape"] for m in metas], None)
    full = lambda: R.get_bboxes(cls, reg, anc, metas, cfg, coder)
    dets = full()
    ms_sel, ms_full = _time(torch, sel, 20), _time(torch, full, 20)
    flt = lambda: R.filter_bboxes(cls, reg, anc, coder, as_batch=True)
    rois = flt()
    ms_flt = _time(torch, flt, 20)
    ms_ref = _time(torch, lambda: R.refine_bboxes(cls, reg1, rois, coder, as_batch=True), 20)
    rows = sum(int(c.size(1) // Cn * c.size(2) * c.size(3)) for c in cls)
    logits_bytes = sum(c.numel() * 4 for c in cls)
    out = {"images": Bn, "rows_per_image": rows, "select_decode_ms": ms_sel, "get_bboxes_ms": ms_full,
           "images_per_s": Bn / ms_full * 1e3, "detections": int(sum(d[0].size(0) for d in dets)),
           "select_decode_gbs": logits_bytes / ms_sel / 1e6, "filter_bboxes_ms": ms_flt, "refine_bboxes_ms": ms_ref,
           "filter_bboxes_gbs": logits_bytes / ms_flt / 1e6}
    if hasattr(R, "get_bboxes_padded"):
        pad = lambda: R.get_bboxes_padded(cls, reg, anc, metas, cfg, coder)
        pad()
        out["get_bboxes_padded_ms"] = _time(torch, pad, 20)
        out["get_bboxes_padded_graph_ms"] = _graph_time(torch, pad)
    return out


def bench_frm(torch, R, dev, hbm):
    """configs[1] FRM shapes: batch 8, 256 channels, 5 FPN levels of a 1024^2 patch; 8 B per element roofline.
    All five levels go through ONE launch sequence (r3g_frm_*_multi_f32), as FeatureRefineModule runs them;
    `per_level_*` is the same work as five separate calls (the reference module's loop).  358 MB in + out per pass: larger
    than L2, no flush needed."""
    from r3det_b200.fr import FrmBackwardPlan, frm_backward, frm_backward_multi, frm_forward, frm_forward_multi
    res = {}
    xs, bts, scales = _frm_inputs(torch, dev, 4)
    elems = sum(x.numel() for x in xs)
    for P in (1, 5):
        tf = _time(torch, lambda: frm_forward_multi(xs, bts, scales, P), 10)
        tb = _time(torch, lambda: frm_backward_multi(xs, bts, scales, P), 10)
        plan = FrmBackwardPlan([tuple(x.shape) for x in xs], bts, scales, P)
        torch.cuda.synchronize()
        tba = _time(torch, lambda: plan.apply(xs), 10)      # the gather alone: what a training step pays when the plan overlapped the forward
        tfl = sum(_time(torch, lambda: frm_forward(x, b, s, P), 10) for x, b, s in zip(xs, bts, scales))
        tbl = sum(_time(torch, lambda: frm_backward(x, b, s, P), 10) for x, b, s in zip(xs, bts, scales))
        res[f"points{P}"] = {"fwd_ms": tf, "bwd_ms": tb, "fwd_gbs": elems * 8 / tf / 1e6, "bwd_gbs": elems * 8 / tb / 1e6,
                             "fwd_frac_of_hbm": elems * 8 / tf / 1e6 / hbm, "bwd_frac_of_hbm": elems * 8 / tb / 1e6 / hbm,
                             "bwd_apply_ms": tba, "bwd_apply_frac_of_hbm": elems * 8 / tba / 1e6 / hbm,
                             "per_level_fwd_ms": tfl, "per_level_bwd_ms": tbl}
    res["elements"] = elems
    res["bytes_per_element"] = 8
    res["roofline"] = {"bound": "hbm", "peak": hbm, "unit": "GB/s", "algorithmic_bytes": elems * 8,
                       "floor_ms": elems * 8 / (hbm * 1e9) * 1e3}
    for k in ("frm_forward_kernel_P1", "frm_backward_kernel_P1"):
        t, src = kernel_traffic(k)
        res["roofline"][k + "_traffic"] = t
    return res


def bench_reference_cuda(torch, dev, line):
    """The reference's own CUDA kernels (oracle/_ref/libref_cuda_*.so: its unmodified .cu files compiled for sm_100) on the
    same B200, same inputs, in the same run — the kernels this library replaces.  ms per call, CUDA events inside the shim."""
    try:
        from oracle import refcuda
    except Exception as e:  # noqa: BLE001
        return {"unavailable": f"{e}"}
    out = {"note": "unmodified reference .cu files behind a C ABI; speed-ups are ours vs these on the same device"}
    gt = torch.from_numpy(rand_obb(GT, 1, "v1")).to(dev)
    an = torch.from_numpy(rand_obb(ANCHORS, 1000, "v1")).to(dev)
    buf = torch.empty((GT, ANCHORS), dtype=torch.float32, device=dev)

    def guard(name, so, fn):
        if not refcuda.available(so):
            out[name] = {"unavailable": so + " not built"}
            return
        try:
            out[name] = fn()
        except Exception as e:  # noqa: BLE001
            out[name] = {"unavailable": f"{type(e).__name__}: {e}"}

    ours_iou = line["iou_variants"]
    guard("v1_iou_1000x200000", "libref_cuda_v1iou.so", lambda: (lambda ms: {
        "ms": ms, "ours_ms": ours_iou["v1_iou"]["ms"], "speedup": ms / ours_iou["v1_iou"]["ms"],
        "kernel": "mat_iou_iof_kernel (rbbox_geo_kernel.cu:230-268)"})(refcuda.v1_iou_ms(gt, an, buf, 3)))
    gt3 = torch.from_numpy(rand_obb(GT, 1, "v3")).to(dev)
    an3 = torch.from_numpy(rand_obb(ANCHORS, 1000, "v3")).to(dev)
    guard("v3_iou_1000x200000", "libref_cuda_v3iou.so", lambda: (lambda ms: {
        "ms": ms, "ours_ms": ours_iou["v3_iou"]["ms"], "speedup": ms / ours_iou["v3_iou"]["ms"],
        "kernel": "box_iou_rotated_cuda_kernel (box_iou_rotated_cuda.cu:13-63)"})(refcuda.v3_iou_ms(gt3, an3, buf, 3)))
    del buf
    # NMS: single class (the reference kernels take no labels; its batched wrappers add class offsets and call the same kernel)
    import r3det_b200 as R
    from r3det_b200._nms_core import nms_device
    for K in (8000, 20000):
        b, s, _ = clustered(K, 2, "v1")
        d6 = torch.from_numpy(np.concatenate([b, s[:, None]], 1)).to(dev)
        Bt, St = d6[:, :5].contiguous(), d6[:, 5].contiguous()
        ours = _time(torch, lambda: nms_device(Bt, St, 0.1, "v1", order_index=True), 10)
        guard(f"v1_nms_K{K}", "libref_cuda_v1nms.so", lambda: (lambda r: {
            "ms": r[0], "kept": r[1], "ours_ms": ours, "speedup": r[0] / ours,
            "kernel": "nmsr_cuda: mask kernel + D2H + host scan (rnms_kernel.cu:229-335)"})(refcuda.v1_nms_ms(d6, 0.1, 2)))
        b3, s3, _ = clustered(K, 2, "v3")
        B3, S3 = torch.from_numpy(b3).to(dev), torch.from_numpy(s3).to(dev)
        ours3 = _time(torch, lambda: nms_device(B3, S3, 0.1, "v3", drop_small=True), 10)
        guard(f"v3_nms_K{K}", "libref_cuda_v3nms.so", lambda: (lambda r: {
            "ms": r[0], "kept": r[1], "ours_ms": ours3, "speedup": r[0] / ours3,
            "kernel": "nms_rotated_cuda (nms_rotated_cuda.cu:71-134)"})(refcuda.v3_nms_ms(B3, S3, 0.1, 2)))
    # polygon NMS
    K = 8000
    b, s, _ = clustered(K, 3, "v1")
    q = torch.cat([R.obb2poly(torch.from_numpy(b).to(dev), "v1"), torch.from_numpy(s).to(dev)[:, None]], 1).contiguous()
    oursp = _time(torch, lambda: R.poly_nms(q, 0.1), 5)
    guard(f"poly_nms_K{K}", "libref_cuda_polynms.so", lambda: (lambda r: {
        "ms": r[0], "kept": r[1], "ours_ms": oursp, "speedup": r[0] / oursp,
        "kernel": "poly_nms_cuda (poly_nms_cuda.cu:196-262)"})(refcuda.poly_nms_ms(q, 0.1, 1)))
    # FRM on the largest level (8 x 256 x 128 x 128), points = 1 (the module's setting) and 5
    from r3det_b200.fr import frm_backward, frm_forward
    xs, bts, scales = _frm_inputs(torch, dev, 4)
    x, bt, sc = xs[0], bts[0], scales[0]
    o = torch.empty_like(x)
    for P in (1, 5):
        of = _time(torch, lambda: frm_forward(x, bt, sc, P), 10)
        ob = _time(torch, lambda: frm_backward(x, bt, sc, P), 10)
        guard(f"frm_fwd_8x256x128x128_P{P}", "libref_cuda_frm.so", lambda: (lambda ms: {
            "ms": ms, "ours_ms": of, "speedup": ms / of, "kernel": "feature_refine_forward_kernel + zero fill (feature_refine_kernel.cu:112-163)"})(
                refcuda.frm_ms(x, bt, sc, P, o, False, 3)))
        guard(f"frm_bwd_8x256x128x128_P{P}", "libref_cuda_frm.so", lambda: (lambda ms: {
            "ms": ms, "ours_ms": ob, "speedup": ms / ob, "kernel": "feature_refine_backward_kernel + zero fill (feature_refine_kernel.cu:165-230)"})(
                refcuda.frm_ms(x, bt, sc, P, o, True, 3)))
    return out


def cpu_nms_baseline(cores):
    """The reference's own CPU NMS (rnms_cpu.cpp:223-282 through oracle/_ref/libref_v1.so; greedy loop, serial) on the clustered
    15-class candidates with batched_rnms's class offsets applied: one image on one thread, and `cores` images on `cores`
    threads (images are independent — the only parallelism the reference's CPU path has)."""
    from concurrent.futures import ThreadPoolExecutor
    try:
        from oracle import ref
        if not ref.available("libref_v1.so"):
            raise FileNotFoundError("oracle/_ref/libref_v1.so")
        fn, kind = (lambda d: ref.v1_nms(d, 0.1)), "reference"
    except Exception:  # noqa: BLE001
        from oracle import port
        fn, kind = (lambda d: port.nms(d[:, :5], d[:, 5], 0.1, "v1", inclusive=True)), "port"
    K = 8000

    def image(seed):
        b, s, l = clustered(K, seed, "v1")
        off = (l.astype(np.float32) * np.float32(b.max() + 1)).astype(np.float32)
        b = b.copy(); b[:, 0] += off; b[:, 1] += off
        return np.ascontiguousarray(np.concatenate([b, s[:, None]], 1), np.float32)

    d0 = image(2)
    t0 = time.perf_counter()
    kept = len(fn(d0))
    t1 = time.perf_counter() - t0
    threads = max(1, cores)
    per_thread = max(1, min(40, int(8.0 / max(t1, 1e-3))))              # ~8 s of wall time
    n_img = threads * per_thread
    imgs = [image(200 + i) for i in range(threads)] * per_thread
    t0 = time.perf_counter()
    with ThreadPoolExecutor(threads) as ex:
        list(ex.map(fn, imgs))
    tn = time.perf_counter() - t0
    return {"value": n_img * K / tn / 1e6, "unit": "Mcands/s", "cores": threads, "kind": kind,
            "single_thread_mcands_per_s": K / t1 / 1e6, "single_thread_ms": t1 * 1e3, "kept": kept,
            "sample": f"K = {K} candidates x 15 classes per image (class offsets as batched_rnms adds them), {n_img} images on {threads} threads, {tn:.1f} s wall"}


def cpu_frm_baseline():
    """FRM has NO CPU implementation in the reference (feature_refine_cuda.cpp is CUDA-only): the CPU column is the plain-C
    oracle restating feature_refine_kernel.cu:112-230 (kind "port"), one thread, on the largest FPN level of the batch (8 x 256 x 128 x 128)."""
    from oracle import port
    rng = np.random.default_rng(4)
    N, Cc, H, stride = 8, 256, 128, 8
    feat = rng.standard_normal((N, Cc, H, H)).astype(np.float32)
    ys_, xs_ = np.meshgrid(np.arange(H) * stride, np.arange(H) * stride, indexing="ij")
    ctr = np.stack([xs_, ys_], -1).reshape(-1, 2).astype(np.float32)
    bx = np.zeros((N, H * H, 5), np.float32)
    bx[:, :, :2] = ctr[None] + rng.normal(0, stride, (N, H * H, 2))
    bx[:, :, 2:4] = np.exp(rng.uniform(np.log(stride), np.log(8 * stride), (N, H * H, 2)))
    bx[:, :, 4] = rng.uniform(-np.pi / 2, 0, (N, H * H))
    bx = bx.reshape(-1, 5)
    t0 = time.perf_counter(); port.frm_forward(feat, bx, 1.0 / stride, 1); tf = time.perf_counter() - t0
    t0 = time.perf_counter(); port.frm_backward(feat, bx, 1.0 / stride, 1); tb = time.perf_counter() - t0
    el = feat.size
    return {"fwd_gbs": el * 8 / tf / 1e9, "bwd_gbs": el * 8 / tb / 1e9, "unit": "GB/s (8 B per element)", "cores": 1, "kind": "port",
            "sample": f"restated oracle (the reference has no CPU FRM): {N} x {Cc} x {H} x {H}, points = 1, fwd {tf:.2f} s / bwd {tb:.2f} s"}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=500)
    ap.add_argument("--warmup", type=int, default=10)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    args = ap.parse_args()
    # stdout carries exactly ONE line (the JSON record): libraries that write there (NCCL prints its version banner on
    # the first communicator) are diverted to stderr for the duration of the run; _emit() restores the descriptor.
    global _STDOUT_FD
    sys.stdout.flush()
    _STDOUT_FD = os.dup(1)
    os.dup2(2, 1)
    try:
        if args.impl == "reference":
            run_reference(args)
        else:
            run_gpu(args)
    finally:
        sys.stdout.flush()
        os.dup2(_STDOUT_FD, 1)


_STDOUT_FD = None


def _emit(record):
    """Print the one JSON line on the real stdout."""
    sys.stdout.flush()
    line = json.dumps(record) + "\n"
    os.write(_STDOUT_FD if _STDOUT_FD is not None else 1, line.encode())


if __name__ == "__main__":
    main()
