#!/usr/bin/env python
"""bench.py — rotated IoU Gpairs/s (headline) + NMS cands/s + FRM GB/s on B200, one JSON line.

    python bench.py [--gpus N] [--steps K] [--warmup W]                  (N > 1: launched by torchrun)
    python bench.py --impl reference [--gpus N] [--steps K] [--warmup W] (the reference's CPU implementation)

A "step" is one pass of the hot path over one batch: BASELINE.json configs[2], the rotated-IoU microbench —
RBboxOverlaps2D_v1 on 1,000 GT x 200,000 anchors per GPU (synthetic rotated boxes), through
the C ABI (r3g_iou_matrix_f32: 2 prep kernels + the pair kernel).  The anchor axis is the shard axis: every
rank owns 200k anchors (weak scaling), GT replicated, no data-path collective.
  value     Gpairs/s, inputs resident in HBM, K steps timed with CUDA events between barriers, max over ranks
  e2e       same metric through the Python plugin API with HOST buffers: pinned H2D of both box sets + D2H of the
            (1000 x 200000) result inside the timed region
  roofline  HBM store bound: 4 B per pair / mean duration of the pair kernel alone (r3g_iou_matrix_prepared_f32),
            against MEASURED_PEAKS.json hbm_gbs
  cpu_baseline  the reference's own host geometry (oracle/_ref, unmodified reference sources) on a bounded sample
Extra objects `nms` and `frm` report the other two kernels of the path with their own rooflines.
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
AR = {'v1': (-np.pi / 2, 0), 'v2': (-np.pi / 4, 3 * np.pi / 4), 'v3': (-np.pi / 2, np.pi / 2)}


# DRAM traffic of one iou_matrix_kernel launch on this workload, from the committed ncu capture (bytes)
NCU_DRAM_BYTES_PER_LAUNCH = 86795776 + 886307840


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
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        d = json.load(open(p))
        return float(d["hbm_gbs"]), "measured (MEASURED_PEAKS.json)"
    return 6650.0, "fallback (B200_PROFILING.md)"


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
        "config": {"workload": "configs[2] rotated IoU microbench: 1000 GT x 200000 anchors per GPU, RBboxOverlaps2D_v1, "
                               "timed on a bounded CPU sample of the same boxes", "variant": VARIANT},
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

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def max_over_ranks(x):
        if world > 1:
            t = torch.tensor([x], dtype=torch.float64, device=dev)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            return float(t.item())
        return x

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
        barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(stream)
        for _ in range(args.steps):
            step()
        e1.record(stream)
        barrier()
        ms_total = max_over_ranks(e0.elapsed_time(e1))
        # the dominant kernel alone (prepared boxes already in the workspace), same K launches
        k0, k1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        k0.record(stream)
        for _ in range(args.steps):
            pair_kernel_only()
        k1.record(stream)
        barrier()
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

    for _ in range(2):
        e2e_step()
    barrier()
    e0.record(stream)
    for _ in range(e2e_steps):
        e2e_step()
    e1.record(stream)
    barrier()
    ms_e2e = max_over_ranks(e0.elapsed_time(e1)) / e2e_steps
    assert float(res_h[0].max()) >= 0.0

    hbm, peak_src = measured_peaks()
    value = world * pairs * args.steps / (ms_total * 1e-3) / 1e9
    achieved = 4.0 * pairs / (ms_kernel * 1e-3) / 1e9
    line = {
        "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": max(args.warmup, 3),
        "ms_per_step": ms_total / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "f32", "data": "synthetic",
        "config": {"workload": "configs[2] rotated IoU microbench: RBboxOverlaps2D_v1, 1000 GT x 200000 anchors per GPU "
                               "(anchor axis sharded across ranks, GT replicated)",
                   "variant": VARIANT, "gt": GT, "anchors_per_gpu": ANCHORS, "strict_reference_parity": True,
                   "l2": "each step writes an 800 MB result (> 126 MB L2); no flush needed"},
        "e2e": {"value": world * pairs / (ms_e2e * 1e-3) / 1e9, "unit": UNIT,
                "h2d_bytes_per_step": int(gt_h.numel() * 4 + an_h.numel() * 4), "d2h_bytes_per_step": int(res_h.numel() * 4),
                "ms_per_step": ms_e2e, "steps": e2e_steps},
        "gpu_launches": 3 * args.steps,
        "roofline": {"bound": "hbm", "achieved": achieved, "peak": hbm, "unit": "GB/s", "frac": achieved / hbm,
                     "traffic": NCU_DRAM_BYTES_PER_LAUNCH, "traffic_source": "ncu --set full, dram__bytes_read.sum + dram__bytes_write.sum "
                     "(profiles/r01_iou_matrix_kernel_ncu_full.txt; 800 MB of it is the result matrix)", "peak_source": peak_src, "kernel": "iou_matrix_kernel<true>",
                     "kernel_ms": ms_kernel, "algorithmic_bytes_per_launch": 4 * pairs,
                     "pairs_circle_pass": stats[0], "pairs_sat_pass": stats[1], "pairs_strict": stats[2]},
        "clocks": clocks.summary(),
        "host_cpus_near_gpu": numa,
    }

    if world > 1:
        line["collectives"] = exercise_collectives(torch, dist, R, dev, rank, world)
    if rank == 0:
        line["iou_variants"] = bench_iou_variants(torch, R, dev)
        line["nms"] = bench_nms(torch, R, dev, hbm)
        line["frm"] = bench_frm(torch, R, dev, hbm)
        line["fused_assign"] = bench_assign(torch, R, dev, gt_h, an_h)
        line["dense_tail"] = bench_dense_tail(torch, R, dev)
        line["train_step_hot_path"] = bench_train_step(torch, R, dev)
        cores = os.cpu_count() or 1
        apt = 12000
        rate, dt, kind = cpu_pairs_per_s(apt, cores)
        line["cpu_baseline"] = {"value": rate / 1e9, "unit": UNIT, "cores": cores, "kind": kind,
                                "sample": f"{GT} GT x {apt * cores} anchors ({apt} per thread, {dt:.1f} s wall)"}
        _emit(line)
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


def exercise_collectives(torch, dist, R, dev, rank, world):
    """N > 1 only: the two harness-level exchanges of the path over NCCL — per-GT assigner statistics
    from the row-sharded IoU (all_reduce MAX of a packed int64 + all_reduce SUM) and ONE all_gather of padded keep
    lists from image-sharded NMS.  Checked for consistency across ranks; timed with CUDA events (not in `value`)."""
    from r3det_b200 import sharding
    gt = torch.from_numpy(rand_obb(GT, 1, VARIANT)).to(dev)
    anchors = torch.from_numpy(rand_obb(ANCHORS, 7, VARIANT)).to(dev)          # same anchors everywhere; each rank takes its rows
    iou_fn = lambda g, a: R.pairwise_iou(g, a, VARIANT)
    e0, ea, e1, e2 = (torch.cuda.Event(enable_timing=True) for _ in range(4))
    local, lo, hi = sharding.sharded_pairwise_iou(gt, anchors, iou_fn)
    torch.cuda.synchronize(); dist.barrier()
    e0.record()
    gmax, garg, npos, nneg = sharding.assigner_stats(local, lo, 0.5, 0.4)
    ea.record(); ea.synchronize()
    stats_ms = e0.elapsed_time(ea)
    images = 8 * world
    ilo, ihi = sharding.shard_range(images, rank, world)
    dets, labels = [], []
    for i in range(ilo, ihi):
        b, s, l = clustered(2000, 50 + i, VARIANT)
        B, S, Lb = (torch.from_numpy(x).to(dev) for x in (b, s, l))
        d, keep = R.batched_rnms(B, S, Lb, 0.1)
        dets.append(d[:2000]); labels.append(Lb[keep][:2000])
    torch.cuda.synchronize(); dist.barrier()
    e1.record()
    all_dets, all_labels = sharding.gather_keep_lists(dets, labels, 2000, images)
    e2.record(); e2.synchronize()
    # every rank must hold the same global view
    chk = torch.tensor([float(gmax.double().sum()), float(garg.double().sum()), float(npos), float(nneg),
                        float(sum(d.double().sum() for d in all_dets))], dtype=torch.float64, device=dev)
    ref = chk.clone(); dist.broadcast(ref, 0)
    ok = bool(torch.equal(chk, ref)) and len(all_dets) == images
    full = R.pairwise_iou(gt, anchors, VARIANT) if rank == 0 else None
    if rank == 0:
        ok = ok and bool(torch.equal(gmax, full.max(dim=1)[0])) and bool(torch.equal(garg, full.max(dim=1)[1]))
    return {"consistent": ok, "images": images, "num_pos": npos, "num_neg": nneg,
            "note": "first call of each collective: includes NCCL communicator warm-up",
            "assigner_stats_first_call_ms": stats_ms, "gather_keep_lists_first_call_ms": e1.elapsed_time(e2)}


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


def bench_iou_variants(torch, R, dev):
    """configs[2] names all three angle conventions: the same 1000 x 200000 matrix for v1 / v2 / v3 (device-resident inputs,
    strict reference parity, whole op = two prepare launches + the pair kernel) and the IoF mode of v1."""
    out = {}
    for v, mode in (("v1", "iou"), ("v2", "iou"), ("v3", "iou"), ("v1", "iof")):
        gt = torch.from_numpy(rand_obb(GT, 1, v)).to(dev)
        an = torch.from_numpy(rand_obb(ANCHORS, 1000, v)).to(dev)
        ms = _time(torch, lambda: R.pairwise_iou(gt, an, v, mode), 30)
        out[f"{v}_{mode}"] = {"ms": ms, "gpairs_per_s": GT * ANCHORS / ms / 1e6}
    return out


def bench_assign(torch, R, dev, gt_h, an_h):
    """the same 1000 x 200000 pairs consumed by the fused MaxIoUAssigner (pos 0.5 / neg 0.4 /
    min_pos 0, gt_max_assign_all): no (G, A) matrix is stored.  `e2e` = pinned host boxes in, assignment (int64 per
    anchor) + max overlaps back on the host."""
    gt, an = gt_h.to(dev), an_h.to(dev)
    fn = lambda: R.max_iou_assign(gt, an, 0.5, 0.4, 0.0, True, True, VARIANT)
    ms = _time(torch, fn, 50)
    res_i = torch.empty((ANCHORS,), dtype=torch.int64).pin_memory()
    res_f = torch.empty((ANCHORS,), dtype=torch.float32).pin_memory()

    def e2e():
        o = R.max_iou_assign(gt_h.to(dev, non_blocking=True), an_h.to(dev, non_blocking=True), 0.5, 0.4, 0.0, True, True, VARIANT)
        res_i.copy_(o.gt_inds, non_blocking=True); res_f.copy_(o.max_overlaps, non_blocking=True)

    ms_e2e = _time(torch, e2e, 20)
    pairs = GT * ANCHORS
    return {"ms": ms, "gpairs_per_s": pairs / ms / 1e6, "e2e_ms": ms_e2e, "e2e_gpairs_per_s": pairs / ms_e2e / 1e6,
            "d2h_bytes_per_step": int(res_i.numel() * 8 + res_f.numel() * 4), "pair_sweeps": 1,
            "num_pos": int((res_i > 0).sum())}


def bench_nms(torch, R, dev, hbm):
    """configs[3]: clustered candidates x 15 DOTA classes per image, nms v1 (batched_rnms semantics), thr 0.1."""
    from r3det_b200._nms_core import nms_device
    out = {"unit": "Mcands/s", "variant": "v1", "classes": 15, "iou_thr": 0.1, "sweep": {}}
    for K in (2000, 8000, 20000, 80000, 200000):
        b, s, l = clustered(K, 2, "v1")
        B, S, Lb = (torch.from_numpy(x).to(dev) for x in (b, s, l))
        scale = torch.tensor(float(b.max() + 1), device=dev)
        fn = lambda: nms_device(B, S, 0.1, "v1", labels=Lb, class_offset=scale, order_index=True)
        keep, num = fn()
        ms = _time(torch, fn, 5 if K >= 80000 else 20)
        out["sweep"][str(K)] = {"ms": ms, "mcands_per_s": K / ms / 1e3, "kept": int(num)}
        if K <= 20000:
            # the same call captured in a CUDA graph (the library never syncs or allocates): device time without the
            # Python wrapper's host overhead, which dominates below ~10k candidates
            try:
                g = torch.cuda.CUDAGraph()
                with torch.cuda.graph(g):
                    fn()
                gms = _time(torch, g.replay, 20)
                out["sweep"][str(K)].update({"graph_ms": gms, "graph_mcands_per_s": K / gms / 1e3})
            except Exception as e:  # noqa: BLE001
                out["sweep"][str(K)]["graph_ms"] = f"capture failed: {e}"
    # configs[3] per-GPU batch: 8 images x K candidates in ONE launch sequence ((image, class) pairs are segments)
    out["batch8"] = {}
    for K in (2000, 8000, 20000):
        imgs = [clustered(K, 100 + i, "v1") for i in range(8)]
        B = torch.from_numpy(np.concatenate([x[0] for x in imgs])).to(dev)
        S = torch.from_numpy(np.concatenate([x[1] for x in imgs])).to(dev)
        Lb = torch.from_numpy(np.concatenate([x[2] for x in imgs])).to(dev)
        bid = torch.arange(8, device=dev).repeat_interleave(K)
        scales = torch.tensor([float(x[0].max() + 1) for x in imgs], device=dev)
        fn = lambda: nms_device(B, S, 0.1, "v1", labels=Lb, class_offset=scales, order_index=True, batch_ids=bid, n_batches=8)
        keep, num = fn()
        ms = _time(torch, fn, 10)
        out["batch8"][str(K)] = {"ms": ms, "mcands_per_s": 8 * K / ms / 1e3, "kept": int(num.sum())}
    return out


def bench_train_step(torch, R, dev):
    """configs[4] / configs[1]: the hot-path part of one R3Det training step on one GPU — 8 patches, per patch the
    assignment of 128 GT against the 196,416 RRetinaNet anchors and against the 21,824 refine-stage boxes (fused
    MaxIoUAssigner, v1), plus FRM forward + backward over the five FPN levels for the batch.  Eager launches and the same
    sequence replayed from a CUDA graph."""
    from r3det_b200.fr import frm_backward_multi, frm_forward_multi
    rng = np.random.default_rng(12)
    gts = [torch.from_numpy(rand_obb(128, 300 + i, "v1", 10, 300)).to(dev) for i in range(8)]
    anc = torch.from_numpy(rand_obb(196416, 400, "v1", 16, 512)).to(dev)
    refs = [torch.from_numpy(rand_obb(21824, 500 + i, "v1", 10, 400)).to(dev) for i in range(8)]
    xs, bts, scales = [], [], []
    for H, stride in ((128, 8), (64, 16), (32, 32), (16, 64), (8, 128)):
        xs.append(torch.randn((8, 256, H, H), device=dev))
        ys_, xs_ = np.meshgrid(np.arange(H) * stride, np.arange(H) * stride, indexing="ij")
        ctr = np.stack([xs_, ys_], -1).reshape(-1, 2).astype(np.float32)
        bx = np.zeros((8, H * H, 5), np.float32)
        bx[:, :, :2] = ctr[None] + rng.normal(0, stride, (8, H * H, 2))
        bx[:, :, 2:4] = np.exp(rng.uniform(np.log(stride), np.log(8 * stride), (8, H * H, 2)))
        bx[:, :, 4] = rng.uniform(-np.pi / 2, 0, (8, H * H))
        bts.append(torch.from_numpy(bx.reshape(-1, 5)).to(dev)); scales.append(1.0 / stride)

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

    ms = _time(torch, step, 10)
    out = {"images": 8, "gt_per_image": 128, "ms": ms, "per_image_assign_calls_ms": _time(torch, step_per_image, 10),
           "pairs": 8 * 128 * (196416 + 21824)}
    try:
        g = torch.cuda.CUDAGraph()
        with torch.cuda.graph(g):
            step()
        out["graph_ms"] = _time(torch, g.replay, 10)
    except Exception as e:  # noqa: BLE001
        out["graph_ms"] = f"capture failed: {e}"
    return out


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
    return {"images": Bn, "rows_per_image": rows, "select_decode_ms": ms_sel, "get_bboxes_ms": ms_full,
            "images_per_s": Bn / ms_full * 1e3, "detections": int(sum(d[0].size(0) for d in dets)),
            "select_decode_gbs": logits_bytes / ms_sel / 1e6, "filter_bboxes_ms": ms_flt, "refine_bboxes_ms": ms_ref,
            "filter_bboxes_gbs": logits_bytes / ms_flt / 1e6}


def bench_frm(torch, R, dev, hbm):
    """configs[1] FRM shapes: batch 8, 256 channels, 5 FPN levels of a 1024^2 patch; 8 B per element roofline.
    All five levels go through ONE launch sequence (r3g_frm_*_multi_f32), as FeatureRefineModule runs them;
    `per_level_*` is the same work as five separate calls (the reference module's loop)."""
    from r3det_b200.fr import FrmBackwardPlan, frm_backward, frm_backward_multi, frm_forward, frm_forward_multi
    rng = np.random.default_rng(4)
    res = {}
    xs, bts, scales = [], [], []
    for H, stride in ((128, 8), (64, 16), (32, 32), (16, 64), (8, 128)):
        xs.append(torch.randn((8, 256, H, H), device=dev))
        ys_, xs_ = np.meshgrid(np.arange(H) * stride, np.arange(H) * stride, indexing="ij")
        ctr = np.stack([xs_, ys_], -1).reshape(-1, 2).astype(np.float32)
        bx = np.zeros((8, H * H, 5), np.float32)
        bx[:, :, :2] = ctr[None] + rng.normal(0, stride, (8, H * H, 2))
        bx[:, :, 2:4] = np.exp(rng.uniform(np.log(stride), np.log(8 * stride), (8, H * H, 2)))
        bx[:, :, 4] = rng.uniform(-np.pi / 2, 0, (8, H * H))
        bts.append(torch.from_numpy(bx.reshape(-1, 5)).to(dev)); scales.append(1.0 / stride)
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
                             "bwd_apply_ms": tba, "per_level_fwd_ms": tfl, "per_level_bwd_ms": tbl}
    res["elements"] = elems
    res["bytes_per_element"] = 8
    return res


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
