#!/usr/bin/env python
"""bench.py -- ray-segments/s of the non-sequential trace (BASELINE.json's metric).

A "step" is one complete trace (all generations) of one batch of synthetic source rays through one of
the BASELINE configs.  Default workload = the north-star run: configs[4], the Michelson
interferometer with a Gaussian source of 1.25e8 gausslets PER GPU (1e9 over the 8 GPUs of a box), far
too large to keep: the source is traced in chunks by `rpx_trace_consume` and every generation is
consumed on the device and freed.  (The 83.5 GB source is 668 MB per generation-chunk row: every
buffer is larger than the 126 MB L2, no flush needed.)

    python bench.py --gpus N --steps K --warmup W            # our CUDA path
    python bench.py --impl reference --steps K --warmup W    # the reference's CPU path

  value     whole-job ray-segments/s of the trace, source resident in HBM (packed ray_t / gausslet_t
            records, as they arrive from the host), device-timed with CUDA events on the engine's stream
            from the first to the last operation of the step (AoS->SoA transposition of every chunk
            included), max over ranks
  e2e       the same metric through the reference-facing call with the source in pinned HOST memory:
            H2D of every chunk (overlapped with the tracing of the previous one) + trace + the
            device-side consumers of the reference's own Michelson example (GaussletCapturePlane at the
            output port -> EFieldPlane: capture filter + E-field summation on a detector grid) + D2H
            of the field and the counts; wall clock, max over ranks
  roofline  dominant kernel (k_shade) achieved algorithmic GB/s vs MEASURED_PEAKS.json (HBM-bound
            workloads), or FP64 instruction rate vs the measured DFMA issue peak (config3)
  cpu_baseline  the reference's own Cython trace (oracle/_ref) on the box's host cores, bounded sample

Smaller workloads (`--workload config2` ...) keep every generation on the device (`rpx_trace_device`)
and ship all of them back in the e2e arm (`rpx_trace_streamed`), as in round 1.

Under torchrun every rank traces its own shard of the source (weak scaling: per-GPU work fixed); no
collective inside the generation loop; per step one NCCL all-gather of the per-generation counts, an
all-reduce of Face.count and (consume mode) an all-reduce of the detector field.
"""
import argparse
import json
import multiprocessing as mp
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

import numpy as np

METRIC = "ray-segments/sec"
WORKLOADS = {
    # the north-star run (BASELINE configs[4]): 1.25e8 gausslets per GPU = 1e9 over 8 GPUs, streamed
    "config5": dict(n=125000000, kw=dict(gausslets=True), mode="consume", block=1 << 21,
                    capture=dict(centre=(0.0, -12.0, 0.0), direction=(0.0, 1.0, 0.0), size=(12.0, 12.0)),
                    detector=dict(y=-14.0, half_width=4.0)),
    # BASELINE configs[3] at its quoted size on one GPU (TIR prism chain, plain rays), streamed
    "config4": dict(n=100000000, kw={}, mode="consume", block=1 << 22, recursion_limit=12, builder="config4_prisms",
                    capture=dict(centre=(0.0, 0.0, 0.0), direction=(1.0, 0.3, 0.0), size=(400.0, 400.0))),
    "config1": dict(n=10000, kw={}),
    "config2": dict(n=1000000, kw={}, capture=dict(centre=(-4.86, -31.3, 0.07), direction=(0.2, 1.0, 0.1), size=(30.0, 30.0))),
    "config3": dict(n=10000000, kw={}),
    "config4_prisms": dict(n=1000000, kw={}, recursion_limit=12),
    "config4_grating": dict(n=1000000, kw={}),
    "config5_1e6": dict(n=1000000, kw=dict(gausslets=True), builder="config5",
                        capture=dict(centre=(0.0, -12.0, 0.0), direction=(0.0, 1.0, 0.0), size=(12.0, 12.0))),
    # triangle-mesh optics: 20480-facet ball lens + 50562-facet mirror, BVH traversal
    "mesh": dict(n=1000000, kw=dict(gausslets=False, ball_subdiv=5, mesh_n=160)),
    # the same optics at STL size: 81920-facet ball lens + 498002-facet mirror
    "mesh_large": dict(n=1000000, kw=dict(gausslets=False, ball_subdiv=6, mesh_n=500), builder="mesh"),
    "config5_rays": dict(n=1000000, kw=dict(gausslets=False), builder="config5",
                         capture=dict(centre=(0.0, -12.0, 0.0), direction=(0.0, 1.0, 0.0), size=(12.0, 12.0))),
}
# algorithmic bytes per ray-segment: read the parent record, write
# back length + end_face_idx, write c children
BYTES_RAY = (188, 12, 188)
BYTES_GAUSSLET = (668, 60, 668)


def workload_cfg(core, name, n, seed):
    from raypier_optics_b200 import configs
    w = WORKLOADS[name]
    key = w.get("builder", name)
    kw = dict(w["kw"])
    kw["n"] = n
    kw["seed"] = seed
    cfg = configs.build(core, key, **kw)
    if "recursion_limit" in w:
        cfg["recursion_limit"] = w["recursion_limit"]
    return cfg


# --------------------------------------------------------------------------- clocks
class ClockSampler(threading.Thread):
    """nvidia-smi clocks / throttle reasons sampled DURING the timed region."""
    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index):
        threading.Thread.__init__(self, daemon=True)
        self.gpu = gpu_index
        self.samples = []
        self.stop_flag = threading.Event()

    def run(self):
        while not self.stop_flag.is_set():
            try:
                out = subprocess.run(["nvidia-smi", "-i", str(self.gpu), "--query-gpu=" + self.Q,
                                      "--format=csv,noheader,nounits"], capture_output=True, text=True,
                                     timeout=5).stdout.strip()
                if out:
                    self.samples.append([x.strip() for x in out.split(",")])
            except Exception:
                pass
            self.stop_flag.wait(0.2)

    def summary(self):
        sm, smax, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for s in self.samples:
            try:
                sm.append(float(s[0]))
                smax.append(float(s[1]))
            except Exception:
                continue
            for name, v in zip(names, s[3:7]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        return {"sm_mhz": float(np.median(sm)) if sm else None,
                "sm_max_mhz": float(max(smax)) if smax else None,
                "reasons": sorted(reasons), "samples": len(sm)}


# --------------------------------------------------------------------------- CPU arms
def _import_ref(flavour):
    from oracle import oracle as O
    return O, O.import_reference(flavour)


def _cpu_worker(args):
    """Trace a shard on one host core with the reference (or the C oracle port)."""
    name, n, seed, kind = args
    from oracle import oracle as O
    if kind == "reference":
        core = O.import_reference("timing") or O.import_reference("parity")
        cfg = workload_cfg(core, name, n, seed)
        rc = O.reference_collection(core, cfg["rays"], cfg["wavelengths"])
        t0 = time.perf_counter()
        traced, _ = O.reference_trace_rays(core, rc, cfg["face_lists"], cfg["recursion_limit"],
                                           cfg["max_length"])
        dt = time.perf_counter() - t0
        return sum(len(t) for t in traced), dt
    import raypier_optics_b200.core as core
    from raypier_optics_b200 import scene
    cfg = workload_cfg(core, name, n, seed)
    sc = scene.Scene(cfg["face_lists"], cfg["wavelengths"])
    t0 = time.perf_counter()
    gens, _ = O.trace_rays(sc, cfg["rays"], cfg["recursion_limit"], cfg["max_length"])
    dt = time.perf_counter() - t0
    return sum(len(g) for g in gens), dt


_POOL = None


def _close_pool():
    global _POOL
    if _POOL is not None:
        _POOL.close()
        _POOL.join()
        _POOL = None


import atexit
atexit.register(_close_pool)


def cpu_pool(cores):
    """One pool of worker processes for the whole run (spawning 16 interpreters per step would cost
    more than the step)."""
    global _POOL
    if _POOL is None and cores > 1:
        _POOL = mp.get_context("spawn").Pool(cores)
    return _POOL


def cpu_trace(name, n_total, cores, kind, seed=1234):
    """Shard n_total source rays over ``cores`` independent worker processes (each with its
    own scene copy; the reference loop is single-threaded under the GIL).  Returns
    (segments, wall seconds including only the trace_rays calls = max over workers)."""
    per = max(n_total // cores, 1)
    jobs = [(name, per, seed + i, kind) for i in range(cores)]
    if cores == 1:
        res = [_cpu_worker(jobs[0])]
    else:
        res = cpu_pool(cores).map(_cpu_worker, jobs)
    segs = sum(r[0] for r in res)
    wall = max(r[1] for r in res)
    return segs, wall


def cpu_kind():
    from oracle import oracle as O
    return "reference" if (O.import_reference("timing") or O.import_reference("parity")) else "port"


def cpu_sample_rays(name, args, cores):
    """Bounded sample of the workload for the CPU arms: ~1-3 s of host work per step."""
    per_core = args.ref_rays_per_core
    if WORKLOADS[name]["kw"].get("gausslets"):
        per_core = max(per_core // 8, 1000)  # 18 segments of 668-byte records per source gausslet
    return min(WORKLOADS[name]["n"], per_core * cores)


def bench_config(name, n, args):
    """The `config` object of the JSON line: static description of the workload, identical for our arm
    and the reference arm (the driver compares them)."""
    w = WORKLOADS[name]
    is_g = bool(w["kw"].get("gausslets"))
    rec = 668 if is_g else 188
    cfg = {"workload": name, "rays_per_gpu": int(n), "record_bytes": rec,
           "record": "gausslet_t" if is_g else "ray_t",
           "mode": w.get("mode", "keep"),
           "parallelism": "source rays sharded by rank, scene replicated",
           "l2": "inputs larger than L2 (%.0f MB per generation)" % (n * rec / 1e6) if n * rec > 126e6
                 else "inputs smaller than L2: not flushed"}
    if w.get("mode") == "consume":
        cfg["chunk_rays"] = int(args.chunk_rays or ((1 << 20) if is_g else (1 << 22)))
        cfg["consumers"] = "capture plane" + (" + %dx%d detector field" % (args.detector_grid, args.detector_grid)
                                              if "detector" in w else "")
        cfg["source"] = "seeded block of %d records repeated to %d" % (min(w["block"], n), n)
    return cfg


def run_reference_arm(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return 0
    kind = cpu_kind()
    cores = os.cpu_count() or 1
    name = args.workload
    n = args.rays if args.rays else WORKLOADS[name]["n"]
    n_sample = cpu_sample_rays(name, args, cores)
    for _ in range(min(args.warmup, 2)):
        cpu_trace(name, max(n_sample // 8, cores), cores, kind)
    segs_total, t_total = 0, 0.0
    for k in range(args.steps):
        segs, wall = cpu_trace(name, n_sample, cores, kind, seed=1000 + k)
        segs_total += segs
        t_total += wall
    value = segs_total / t_total
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": "ray-segments/s",
        "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": 1e3 * t_total / max(args.steps, 1), "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": bench_config(name, n, args),
        "note": "reference Cython trace_rays on host cores, source sharded over independent worker "
                "processes (the reference loop is single-threaded); each step a bounded sample of the workload",
        "cpu_baseline": {"value": value, "unit": "ray-segments/s", "cores": cores, "kind": kind,
                         "sample": "%d source rays of %s per step, %d steps (the rate does not depend on the "
                                   "source size: rays are independent)" % (n_sample, name, args.steps)},
        "e2e": {"value": value, "unit": "ray-segments/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line))
    return 0


# --------------------------------------------------------------------------- our arm
def load_peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        try:
            d = json.load(open(p))
            return float(d["hbm_gbs"]), "measured (MEASURED_PEAKS.json)"
        except Exception:
            pass
    return 6650.0, "fallback (B200_PROFILING.md)"


FP64_PEAK_TINST = 63.6 * 148 * 1.965e9 / 1e12  # T fp64 lane-instructions/s, measured (profiles/microbench/fp64_latency.cu)


def load_fp64(workload):
    """fp64 instruction counts per launch of the FP64-bound workloads, from the committed ncu capture."""
    p = os.path.join(ROOT, "profiles", "roofline_latest.json")
    try:
        return json.load(open(p)).get("fp64", {}).get(workload)
    except Exception:
        return None


def load_traffic(workload, kernel):
    """dram bytes per launch of the dominant kernel from the committed ncu capture."""
    p = os.path.join(ROOT, "profiles", "roofline_latest.json")
    try:
        d = json.load(open(p))
        e = d.get(workload, {}).get(kernel)
        return e
    except Exception:
        return None


def run_ours(args):
    import torch
    import torch.distributed as dist
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    distributed = world > 1
    if not torch.cuda.is_available():
        raise RuntimeError("bench.py needs a CUDA device: raypier_optics_b200 has no CPU fallback")
    torch.cuda.set_device(local_rank)
    if distributed:
        dist.init_process_group(backend="nccl", device_id=torch.device("cuda", local_rank))
    dev = torch.device("cuda", local_rank)

    import raypier_optics_b200.core as core
    from raypier_optics_b200 import scene
    from raypier_optics_b200 import distributed as rdist
    from raypier_optics_b200.engine import Engine

    name = args.workload
    n = args.rays if args.rays else WORKLOADS[name]["n"]
    cfg = workload_cfg(core, name, n, seed=100 + rank)  # every rank traces its own shard
    sc = scene.Scene(cfg["face_lists"], cfg["wavelengths"])
    eng = Engine(local_rank)
    eng.set_scene(sc)
    rays = cfg["rays"]
    is_g = rays.dtype.itemsize == 668
    ml, rl = cfg["max_length"], cfg["recursion_limit"]

    def barrier():
        torch.cuda.synchronize()
        if distributed:
            dist.barrier()
        torch.cuda.synchronize()

    # ---------------- device-resident arm (`value`)
    pristine = eng.upload(rays)

    def step_resident():
        d = eng.clone(pristine)
        res = eng.trace_device(d, ml, rl)
        if distributed:  # the only collectives of a trace: counts all-gather + Face.count all-reduce
            rdist.exchange_counts(res.counts, res.face_counts, device=dev)
        return res

    for _ in range(args.warmup):
        step_resident().free()
    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()
    barrier()
    t0 = time.perf_counter()
    dev_ms, segs, launches = 0.0, 0, 0
    k_ms = {"intersect": [0.0, 0], "shade": [0.0, 0]}
    gen_counts = None
    for _ in range(args.steps):
        res = step_resident()
        dev_ms += res.device_ms
        segs += res.segments
        launches += res.launches
        for k in k_ms:
            k_ms[k][0] += res.kernel_ms[k][0]
            k_ms[k][1] += res.kernel_ms[k][1]
        gen_counts = res.counts
        res.free()
    barrier()
    wall_ms = 1e3 * (time.perf_counter() - t0)
    if rank == 0:
        sampler.stop_flag.set()
        sampler.join()

    # max over ranks of the device time, sum over ranks of the work
    t = torch.tensor([dev_ms, wall_ms], dtype=torch.float64, device=dev)
    s = torch.tensor([segs, launches], dtype=torch.float64, device=dev)
    if distributed:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        dist.all_reduce(s, op=dist.ReduceOp.SUM)
    dev_ms_max, wall_ms_max = float(t[0]), float(t[1])
    segs_all, launches_all = float(s[0]), float(s[1])
    value = segs_all / (dev_ms_max * 1e-3)

    # ---------------- end-to-end arm (`e2e`): host buffers, copies inside the timed region
    rec = rays.dtype.itemsize
    pinned_in = eng.pinned_empty(rays.shape[0], rays.dtype)
    pinned_in[:] = rays
    # output staging sized from the known generation profile (+ slack)
    out_cap = int(sum(gen_counts) * 1.05) + 1024
    pinned_out = eng.pinned_empty(out_cap, rays.dtype)
    pinned_out2 = eng.pinned_empty(int(sum(gen_counts) * 1.05) + 1024 * (len(gen_counts) + 2), rays.dtype)

    def step_e2e_oneshot():
        res = eng.trace(pinned_in, ml, rl)
        off = 0
        for g in range(res.n_generations):
            c = res.counts[g]
            res.generation(g, out=pinned_out[off:off + c])
            off += c
        if distributed:
            rdist.exchange_counts(res.counts, res.face_counts, device=dev)
        nseg = res.segments
        res.free()
        return nseg, off

    # the public streamed call: source cut into chunks, H2D of chunk c+1 and D2H of chunk c's
    # generations overlap the tracing (full-duplex PCIe); byte-identical results
    # (tests/test_parity_gpu.py::test_streamed_trace_is_identical_to_one_shot)
    gen_caps = [int(c * 1.05) + 1024 for c in gen_counts] + [1024]
    outs, off = [], 0
    for c in gen_caps:
        outs.append(pinned_out2[off:off + c])
        off += c

    def step_e2e():
        gens, fc, _ = eng.trace_streamed(pinned_in, ml, rl, outs, chunk_rays=args.chunk_rays)
        counts = [len(g) for g in gens]
        if distributed:
            rdist.exchange_counts(counts, fc, device=dev)
        return sum(counts), sum(counts)

    e2e_steps = max(1, min(args.steps, args.e2e_steps))
    for _ in range(min(args.warmup, 2)):
        step_e2e()
    barrier()
    t0 = time.perf_counter()
    e2e_segs, d2h_records = 0, 0
    for _ in range(e2e_steps):
        a, b = step_e2e()
        e2e_segs += a
        d2h_records = b
    barrier()
    e2e_s = time.perf_counter() - t0
    te = torch.tensor([e2e_s], dtype=torch.float64, device=dev)
    se = torch.tensor([e2e_segs], dtype=torch.float64, device=dev)
    if distributed:
        dist.all_reduce(te, op=dist.ReduceOp.MAX)
        dist.all_reduce(se, op=dist.ReduceOp.SUM)
    e2e_value = float(se[0]) / float(te[0])
    # the same through the one-shot call (upload, trace, then download generation by generation)
    step_e2e_oneshot()
    barrier()
    t0 = time.perf_counter()
    one_segs = 0
    for _ in range(e2e_steps):
        one_segs += step_e2e_oneshot()[0]
    barrier()
    e2e_oneshot = one_segs / (time.perf_counter() - t0)

    # ---------------- optional capture-plane arm (--capture): the same e2e call, but
    # the generations stay on the device, are filtered there (rpx_capture) and only the captured
    # rays cross PCIe -- what a RayCapturePlane downstream of the trace actually needs
    capture = None
    if args.capture and "capture" in WORKLOADS[name]:
        from raypier_optics_b200 import configs
        cp = WORKLOADS[name]["capture"]
        face = core.cfaces.RectangularFace(length=cp["size"][0], width=cp["size"][1], offset=0.0, z_plane=0.0)
        fl = core.ctracer.FaceList(owner=configs.Pose(centre=cp["centre"], direction=cp["direction"]))
        fl.faces = [face]
        fl.sync_transforms()
        eng.set_capture_scene(scene.Scene([fl], np.asarray([1.0])))

        def step_capture():
            res = eng.trace(pinned_in, ml, rl)
            got, _, counts = res.capture(cfg["wavelengths"], out=pinned_out)
            nseg = res.segments
            res.free()
            return nseg, len(got), counts

        for _ in range(2):
            step_capture()
        barrier()
        t0 = time.perf_counter()
        cap_segs = 0
        for _ in range(e2e_steps):
            a, ncap, cap_counts = step_capture()
            cap_segs += a
        barrier()
        cap_s = time.perf_counter() - t0
        capture = {"value": cap_segs / cap_s, "unit": "ray-segments/s (this rank)", "captured_rays": int(ncap),
                   "captured_per_generation": cap_counts, "h2d_bytes_per_step": int(rays.shape[0] * rec),
                   "d2h_bytes_per_step": int(ncap * rec), "plane": cp}

    if rank != 0:
        if distributed:
            dist.destroy_process_group()
        return 0

    # ---------------- roofline of the dominant kernel
    peak, peak_src = load_peaks()
    b_in, b_wb, b_child = BYTES_GAUSSLET if is_g else BYTES_RAY
    per_step_parents = sum(gen_counts)
    per_step_children = sum(gen_counts[1:])  # children actually emitted and traced
    # children emitted by the last traced generation are built too (then found empty / dropped)
    shade_ms_avg = k_ms["shade"][0] / max(k_ms["shade"][1], 1)
    isect_ms_avg = k_ms["intersect"][0] / max(k_ms["intersect"][1], 1)
    n_launch = max(len(gen_counts), 1)
    shade_bytes_per_launch = (b_in * per_step_parents + b_child * per_step_children) / n_launch
    isect_bytes_per_launch = ((48 if not is_g else 48) + b_wb) * per_step_parents / n_launch
    dominant = "k_shade" if k_ms["shade"][0] >= k_ms["intersect"][0] else "k_intersect"
    if dominant == "k_shade":
        ach = shade_bytes_per_launch / (shade_ms_avg * 1e-3) / 1e9
    else:
        ach = isect_bytes_per_launch / (isect_ms_avg * 1e-3) / 1e9
    gen_bytes = (b_in + b_wb) * per_step_parents + b_child * per_step_children
    gen_ach = gen_bytes / ((k_ms["shade"][0] + k_ms["intersect"][0]) / args.steps * 1e-3) / 1e9
    traffic = load_traffic(name, dominant)
    roofline = {"bound": "hbm", "kernel": dominant, "achieved": ach, "peak": peak, "unit": "GB/s",
                "frac": ach / peak, "traffic": traffic, "peak_source": peak_src,
                "per_launch_ms": {"k_shade": shade_ms_avg, "k_intersect": isect_ms_avg},
                "generation": {"achieved": gen_ach, "frac": gen_ach / peak,
                               "note": "388 B/segment model over k_intersect + k_shade together"}}
    fp64 = load_fp64(name)
    if fp64 is not None:
        # FP64-bound workload (config3: Newton on the asphere, secant + Zernike tape on the distorted face):
        # fp64 instructions per ray counted by ncu (smsp__sass_thread_inst_executed_op_d{add,mul,fma}_pred_on,
        # profiles/roofline_latest.json) x rays / kernel time, against the measured DFMA issue peak
        n0 = float(gen_counts[0])
        k_time = {"k_shade": shade_ms_avg, "k_intersect": isect_ms_avg}
        k_step_ms = {"k_shade": k_ms["shade"][0] / args.steps, "k_intersect": k_ms["intersect"][0] / args.steps}
        per = {}
        for kname in ("k_intersect", "k_shade"):
            e = fp64.get(kname)  # instruction totals over ALL launches of that kernel in one trace of e["rays"] source rays
            if e and k_step_ms[kname] > 0:
                inst = (e["dadd"] + e["dmul"] + e["dfma"]) / e["rays"] * n0
                flop = (e["dadd"] + e["dmul"] + 2 * e["dfma"]) / e["rays"] * n0
                per[kname] = {"fp64_inst_per_source_ray": inst / n0, "flop_per_source_ray": flop / n0,
                              "achieved_tinst_s": inst / (k_step_ms[kname] * 1e-3) / 1e12,
                              "achieved_tflops": flop / (k_step_ms[kname] * 1e-3) / 1e12,
                              "fp64_pipe_pct_ncu": e.get("fp64_pipe_pct")}
        if dominant in per:
            peak_inst = FP64_PEAK_TINST
            roofline = {"bound": "fp64", "kernel": dominant, "achieved": per[dominant]["achieved_tinst_s"], "peak": peak_inst,
                        "unit": "T fp64 inst/s", "frac": per[dominant]["achieved_tinst_s"] / peak_inst,
                        "traffic": traffic, "peak_source": "measured DFMA issue rate, 63.6 lanes/clk/SM x 148 SMs x 1.965 GHz "
                        "(profiles/microbench/fp64_latency.cu) = 36.4 TFLOP/s",
                        "tflops": per[dominant]["achieved_tflops"], "tflops_peak": 2 * peak_inst,
                        "per_kernel": per, "per_launch_ms": k_time,
                        "hbm": {"achieved": ach, "peak": peak, "frac": ach / peak, "unit": "GB/s"}}

    # ---------------- CPU baseline (rank 0, N=1 only, bounded sample)
    cpu = None
    if world == 1 and not args.no_cpu_baseline:
        kind = cpu_kind()
        cores = os.cpu_count() or 1
        n_sample = cpu_sample_rays(name, args, cores)
        csegs, cwall = cpu_trace(name, n_sample, cores, kind)
        cpu = {"value": csegs / cwall, "unit": "ray-segments/s", "cores": cores, "kind": kind,
               "sample": "%d source rays of %s sharded over %d worker processes" % (n_sample, name, cores)}

    line = {
        "metric": METRIC, "value": value, "unit": "ray-segments/s", "n_gpus": world,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": dev_ms_max / args.steps,
        "wall_ms_per_step": wall_ms_max / args.steps, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": bench_config(name, int(rays.shape[0]), args),
        "trace": {"generations": gen_counts, "segments_per_step_per_gpu": int(per_step_parents)},
        "e2e": {"value": e2e_value, "unit": "ray-segments/s",
                "h2d_bytes_per_step": int(rays.shape[0] * rec),
                "d2h_bytes_per_step": int(d2h_records * rec),
                "steps": e2e_steps, "api": "Engine.trace_streamed (rpx_trace_streamed), chunk %d rays" % args.chunk_rays,
                "one_shot_value_this_rank": e2e_oneshot},
        "gpu_launches": int(launches_all),
        "clocks": sampler.summary(),
        "roofline": roofline,
        "cpu_baseline": cpu,
    }
    if capture is not None:
        line["e2e_capture"] = capture
    print(json.dumps(line))
    if distributed:
        dist.destroy_process_group()
    return 0


def _mem_available_bytes():
    """Host memory this process may count on: MemAvailable, capped by the container's cgroup limit."""
    avail = 64 << 30
    try:
        for line in open("/proc/meminfo"):
            if line.startswith("MemAvailable:"):
                avail = int(line.split()[1]) * 1024
    except Exception:
        pass
    for path, used in (("/sys/fs/cgroup/memory.max", "/sys/fs/cgroup/memory.current"),
                       ("/sys/fs/cgroup/memory/memory.limit_in_bytes", "/sys/fs/cgroup/memory/memory.usage_in_bytes")):
        try:
            lim = open(path).read().strip()
            if lim.isdigit() and int(lim) < (1 << 60):
                cur = int(open(used).read().strip())
                avail = min(avail, max(int(lim) - cur, 0))
        except Exception:
            pass
    return avail


def _fill_repeating(dst_u8, block_u8, threads=8):
    """dst := block repeated (numpy releases the GIL in copies: a few threads saturate host DRAM)."""
    import concurrent.futures as cf
    B, n = block_u8.shape[0], dst_u8.shape[0]
    jobs = [(lo, min(lo + B, n)) for lo in range(0, n, B)]

    def one(j):
        lo, hi = j
        dst_u8[lo:hi] = block_u8[:hi - lo]

    with cf.ThreadPoolExecutor(threads) as ex:
        list(ex.map(one, jobs))


def run_consume(args):
    """The north-star workloads: sources too large to keep (1e8 rays / 1.25e8 gausslets per GPU), traced in
    chunks by rpx_trace_consume with device-side consumers."""
    import torch
    import torch.distributed as dist
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    local_world = int(os.environ.get("LOCAL_WORLD_SIZE", str(world)))
    distributed = world > 1
    if not torch.cuda.is_available():
        raise RuntimeError("bench.py needs a CUDA device: raypier_optics_b200 has no CPU fallback")
    torch.cuda.set_device(local_rank)
    if distributed:
        dist.init_process_group(backend="nccl", device_id=torch.device("cuda", local_rank))
    dev = torch.device("cuda", local_rank)

    import raypier_optics_b200.core as core
    from raypier_optics_b200 import configs, scene
    from raypier_optics_b200 import distributed as rdist
    from raypier_optics_b200.engine import Engine

    numa_cores = rdist.bind_to_gpu_numa(local_rank)  # before any pinned allocation (first touch)
    name = args.workload
    w = WORKLOADS[name]
    n_req = args.rays if args.rays else w["n"]
    is_g = bool(w["kw"].get("gausslets"))
    rec = 668 if is_g else 188
    chunk = int(args.chunk_rays or ((1 << 20) if is_g else (1 << 22)))
    # HBM budget: the resident source + the working set of two chunks (bound-sized generation buffers of the
    # pipelined loop, ~64 chunk-sized buffers for a branching trace, + the capture output)
    free_b, _total_b = torch.cuda.mem_get_info(dev)
    working = 80 * chunk * rec + (2 << 30)
    n = int(min(n_req, max((free_b * 0.94 - working) // rec, chunk)))
    block_n = int(min(w["block"], n))
    t_setup = time.perf_counter()
    cfg = workload_cfg(core, name, block_n, seed=100 + rank)  # every rank traces its own shard
    block = np.ascontiguousarray(cfg["rays"])
    assert block.dtype.itemsize == rec
    sc = scene.Scene(cfg["face_lists"], cfg["wavelengths"])
    eng = Engine(local_rank)
    eng.set_scene(sc)
    ml, rl = cfg["max_length"], cfg["recursion_limit"]
    cp = w["capture"]
    face = core.cfaces.RectangularFace(length=cp["size"][0], width=cp["size"][1], offset=0.0, z_plane=0.0)
    fl = core.ctracer.FaceList(owner=configs.Pose(centre=cp["centre"], direction=cp["direction"]))
    fl.faces = [face]
    fl.sync_transforms()
    eng.set_capture_scene(scene.Scene([fl], np.asarray([1.0])))
    det, npt = None, 0
    if "detector" in w:
        side = args.detector_grid
        xs = np.linspace(-w["detector"]["half_width"], w["detector"]["half_width"], side)
        gx, gz = np.meshgrid(xs, xs)
        pts = np.ascontiguousarray(np.stack([gx.ravel(), np.full(gx.size, w["detector"]["y"]), gz.ravel()], axis=1))
        det = eng.detector(pts, cfg["wavelengths"])
        npt = len(pts)

    def barrier():
        torch.cuda.synchronize()
        if distributed:
            dist.barrier()
        torch.cuda.synchronize()

    # ---------------- the source, resident in HBM as packed records (the seeded block repeated)
    blk_u8 = torch.from_numpy(block.view(np.uint8).reshape(-1)).to(dev)
    src = torch.empty(n * rec, dtype=torch.uint8, device=dev)
    B = blk_u8.numel()
    reps = (n * rec) // B
    if reps:
        src[:reps * B].view(reps, B).copy_(blk_u8.unsqueeze(0).expand(reps, B))
    if n * rec > reps * B:
        src[reps * B:].copy_(blk_u8[:n * rec - reps * B])
    torch.cuda.synchronize()
    del blk_u8

    def collectives(counts, fc, with_field):
        if not distributed:
            return
        rdist.exchange_counts(counts, fc, device=dev)
        if with_field and det is not None:
            rdist.allreduce_detector(eng, det)

    def step_resident(consumers):
        if det is not None and consumers:
            det.reset()
        r = eng.trace_consume(src.data_ptr(), ml, rl, n=n, is_gausslet=is_g, chunk_rays=chunk, capture=consumers,
                              detector=det if consumers else None)
        collectives(r.counts, r.face_counts, consumers)
        return r

    for _ in range(args.warmup):
        step_resident(False)
    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()
    barrier()
    t0 = time.perf_counter()
    dev_ms, trace_ms, segs, launches = 0.0, 0.0, 0, 0
    k_ms = {"intersect": [0.0, 0], "shade": [0.0, 0]}
    gen_counts = None
    for _ in range(args.steps):
        r = step_resident(False)
        dev_ms += r.device_ms
        trace_ms += r.trace_ms
        segs += r.segments
        launches += r.launches
        for k in k_ms:
            k_ms[k][0] += r.kernel_ms[k][0]
            k_ms[k][1] += r.kernel_ms[k][1]
        gen_counts = r.counts
    barrier()
    wall_ms = 1e3 * (time.perf_counter() - t0)
    if rank == 0:
        sampler.stop_flag.set()
        sampler.join()
    t = torch.tensor([dev_ms, wall_ms, trace_ms], dtype=torch.float64, device=dev)
    sm = torch.tensor([segs, launches], dtype=torch.float64, device=dev)
    if distributed:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        dist.all_reduce(sm, op=dist.ReduceOp.SUM)
    dev_ms_max, wall_ms_max, trace_ms_max = float(t[0]), float(t[1]), float(t[2])
    segs_all, launches_all = float(sm[0]), float(sm[1])
    value = segs_all / (dev_ms_max * 1e-3)

    # the same, consumers on (capture plane + detector field), device-timed: one step
    step_resident(True)
    barrier()
    rc = step_resident(True)
    barrier()
    with_consumers = {"value_this_rank": rc.segments / (rc.device_ms * 1e-3), "device_ms": rc.device_ms,
                      "captured_rays": rc.n_captured, "detector_ms": det.ms if det is not None else None,
                      "detector_points": npt}
    del src
    torch.cuda.empty_cache()

    # ---------------- end-to-end arm: source in pinned HOST memory, every chunk crosses PCIe
    avail = _mem_available_bytes()
    # default: at most 32 GB of page-locked source per rank (8 ranks page-locking 83.5 GB each would spend minutes
    # of set-up in the kernel's page pinning); a smaller host source is streamed in several passes per step --
    # the bytes crossing PCIe per step are the same, `host_source_rays` / `passes_per_step` say what was done
    budget = args.host_source_gb * (1 << 30) if args.host_source_gb else min(0.45 * avail / max(local_world, 1), 32 << 30)
    n_host = int(max(min(n, budget // rec), min(n, chunk)))
    if chunk <= n_host < n:
        n_host = n_host // chunk * chunk  # whole chunks per pass
    pinned = None
    while pinned is None:
        try:
            pinned = eng.pinned_empty(n_host, block.dtype)
        except MemoryError:  # page-locking refused (ulimit / cgroup): stream a smaller host source in more passes
            if n_host <= chunk:
                raise
            n_host = max(n_host // 2, min(n, chunk))
    _fill_repeating(pinned.view(np.uint8).reshape(-1), block.view(np.uint8).reshape(-1))
    passes = [(lo, min(lo + n_host, n)) for lo in range(0, n, n_host)]
    setup_s = time.perf_counter() - t_setup

    def step_e2e(consumers=True):
        if det is not None and consumers:
            det.reset()
        nseg, ncap, counts, fc = 0, 0, None, None
        for lo, hi in passes:
            r = eng.trace_consume(pinned[:hi - lo], ml, rl, chunk_rays=chunk, capture=consumers,
                                  detector=det if consumers else None)
            nseg += r.segments
            ncap += r.n_captured
            counts = r.counts if counts is None else [a + b for a, b in zip(counts, r.counts)]
            fc = r.face_counts if fc is None else fc + r.face_counts
        collectives(counts, fc, consumers)
        E = det.read() if (det is not None and consumers) else None   # D2H of the step's result
        return nseg, ncap, E

    e2e_steps = max(1, min(args.steps, args.e2e_steps))
    step_e2e()
    barrier()
    t0 = time.perf_counter()
    e2e_segs = 0
    for _ in range(e2e_steps):
        a, ncap, E = step_e2e()
        e2e_segs += a
    barrier()
    e2e_s = time.perf_counter() - t0
    # trace only (no consumers): H2D + trace + D2H of the counts
    barrier()
    t0 = time.perf_counter()
    a_only, _, _ = step_e2e(False)
    barrier()
    e2e_only_s = time.perf_counter() - t0
    te = torch.tensor([e2e_s, e2e_only_s], dtype=torch.float64, device=dev)
    se = torch.tensor([e2e_segs, a_only], dtype=torch.float64, device=dev)
    if distributed:
        dist.all_reduce(te, op=dist.ReduceOp.MAX)
        dist.all_reduce(se, op=dist.ReduceOp.SUM)
    e2e_value = float(se[0]) / float(te[0])
    e2e_only_value = float(se[1]) / float(te[1])
    field_power = float((np.abs(E) ** 2).sum()) if E is not None else None

    if rank != 0:
        if distributed:
            dist.destroy_process_group()
        return 0

    # ---------------- roofline of the dominant kernel
    peak, peak_src = load_peaks()
    b_in, b_wb, b_child = BYTES_GAUSSLET if is_g else BYTES_RAY
    parents = sum(gen_counts)
    children = sum(gen_counts[1:])
    shade_ms_avg = k_ms["shade"][0] / max(k_ms["shade"][1], 1)
    isect_ms_avg = k_ms["intersect"][0] / max(k_ms["intersect"][1], 1)
    shade_launches_per_step = k_ms["shade"][1] / args.steps
    shade_bytes_per_launch = (b_in * parents + b_child * children) / max(shade_launches_per_step, 1)
    ach = shade_bytes_per_launch / (shade_ms_avg * 1e-3) / 1e9
    gen_bytes = (b_in + b_wb) * parents + b_child * children
    gen_ach = gen_bytes / ((k_ms["shade"][0] + k_ms["intersect"][0]) / args.steps * 1e-3) / 1e9
    step_ach = gen_bytes / (dev_ms_max / args.steps * 1e-3) / 1e9
    roofline = {"bound": "hbm", "kernel": "k_shade", "achieved": ach, "peak": peak, "unit": "GB/s", "frac": ach / peak,
                "traffic": load_traffic(name, "k_shade"), "peak_source": peak_src,
                "units_per_launch": parents / max(shade_launches_per_step, 1),
                "per_launch_ms": {"k_shade": shade_ms_avg, "k_intersect": isect_ms_avg},
                "kernel_share_of_step": {"k_shade": k_ms["shade"][0] / (dev_ms if dev_ms else 1),
                                         "k_intersect": k_ms["intersect"][0] / (dev_ms if dev_ms else 1)},
                "generation": {"achieved": gen_ach, "frac": gen_ach / peak,
                               "note": "%d B/segment model over k_intersect + k_shade together" % (b_in + b_wb + b_child)},
                "step": {"achieved": step_ach, "frac": step_ach / peak,
                         "note": "same bytes over the whole device-timed step (transposition of every chunk included)"}}

    # ---------------- the Python drop-in itself (core.tracer.trace_rays with collection objects in and out:
    # flatten + upload + trace + download of EVERY generation + container objects), bounded size
    from raypier_optics_b200.core import tracer as T
    n_drop = int(min(args.dropin_rays, block_n))
    cls = core.ctracer.GaussletCollection if is_g else core.ctracer.RayCollection

    rc_ = cls.from_array(block[:n_drop])  # the caller's source collection, re-traced like a model does on every change
    rc_.wavelengths = cfg["wavelengths"]

    def step_dropin():
        traced, _ = T.trace_rays(rc_, cfg["face_lists"], recursion_limit=rl, max_length=ml, device=local_rank)
        assert traced[0] is rc_
        return sum(len(t_) for t_ in traced)

    t0 = time.perf_counter()
    step_dropin()  # first call of the process: page-locks the result blocks (raypier_optics_b200/_hostpool.py)
    first_s = time.perf_counter() - t0
    t0 = time.perf_counter()
    drop_segs = step_dropin() + step_dropin()
    from raypier_optics_b200._hostpool import get_pool
    dropin = {"value": drop_segs / (time.perf_counter() - t0), "unit": "ray-segments/s (this rank)", "rays": n_drop,
              "first_call_value": drop_segs / 2 / first_s, "result_pool": get_pool(None).stats(),
              "api": "raypier_optics_b200.core.tracer.trace_rays (collections in, list of collections out; every "
                     "generation crosses PCIe into pooled page-locked result arrays, the source through pageable memory)"}

    cpu = None
    if world == 1 and not args.no_cpu_baseline:
        kind = cpu_kind()
        cores = os.cpu_count() or 1
        n_sample = cpu_sample_rays(name, args, cores)
        csegs, cwall = cpu_trace(name, n_sample, cores, kind)
        cpu = {"value": csegs / cwall, "unit": "ray-segments/s", "cores": cores, "kind": kind,
               "sample": "%d source rays of %s sharded over %d worker processes (the rate does not depend on the "
                         "source size: rays are independent; the full workload is %d x this sample)"
                         % (n_sample, name, cores, max(n // max(n_sample, 1), 1))}

    config = bench_config(name, n, args)
    if n != n_req:
        config["rays_per_gpu_requested"] = int(n_req)
    line = {
        "metric": METRIC, "value": value, "unit": "ray-segments/s", "n_gpus": world,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": dev_ms_max / args.steps,
        "wall_ms_per_step": wall_ms_max / args.steps, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": config,
        "trace": {"generations": gen_counts, "segments_per_step_per_gpu": int(parents), "chunks_per_step": int(rc.n_chunks),
                  "generation_loop_ms_per_step": trace_ms_max / args.steps, "setup_s": setup_s},
        "value_with_consumers": with_consumers,
        "e2e_dropin": dropin,
        "e2e": {"value": e2e_value, "unit": "ray-segments/s",
                "h2d_bytes_per_step": int(n * rec),
                "d2h_bytes_per_step": int(npt * 48 + 8 * len(gen_counts) + 4 * sc.n_traced_faces),
                "steps": e2e_steps, "captured_rays_per_step": int(ncap), "detector_points": npt,
                "field_power": field_power,
                "host_source_rays": n_host, "passes_per_step": len(passes),
                "numa_bound_cores": len(numa_cores) if numa_cores else None,
                "api": "Engine.trace_consume (rpx_trace_consume): pinned host source, chunk %d, capture plane%s, "
                       "D2H = field + counts" % (chunk, " + detector field" if det is not None else ""),
                "trace_only_value": e2e_only_value},
        "gpu_launches": int(launches_all),
        "clocks": sampler.summary(),
        "roofline": roofline,
        "cpu_baseline": cpu,
    }
    print(json.dumps(line))
    if distributed:
        dist.destroy_process_group()
    return 0


def run_fields(args):
    """--fields: the E-field summation (cfields.sum_gaussian_modes) as its own
    measurement.  Workload: the Michelson output gausslets (config 5 traced on the GPU, rays
    captured at the output port) summed on a square detector grid.  One "step" = one evaluation of
    all N_ray x N_pt pairs.  Prints one JSON line (metric: mode-point pairs/s)."""
    import torch
    if not torch.cuda.is_available():
        raise RuntimeError("bench.py needs a CUDA device: raypier_optics_b200 has no CPU fallback")
    import raypier_optics_b200.core as core
    from raypier_optics_b200 import configs, scene
    from raypier_optics_b200.engine import Engine
    eng = Engine(0)
    n_src = args.rays if args.rays else 50000
    cfg = configs.build(core, "config5", n=n_src, gausslets=True, seed=7)
    eng.set_scene(scene.Scene(cfg["face_lists"], cfg["wavelengths"]))
    face = core.cfaces.RectangularFace(length=12.0, width=12.0, offset=0.0, z_plane=0.0)
    fl = core.ctracer.FaceList(owner=configs.Pose(centre=(0.0, -12.0, 0.0), direction=(0.0, 1.0, 0.0)))
    fl.faces = [face]
    fl.sync_transforms()
    eng.set_capture_scene(scene.Scene([fl], np.asarray([1.0])))
    res = eng.trace(np.ascontiguousarray(cfg["rays"]), cfg["max_length"], cfg["recursion_limit"])
    g, _, _ = res.capture(cfg["wavelengths"])
    res.free()
    side = args.field_grid
    xs = np.linspace(-4.0, 4.0, side)
    gx, gz = np.meshgrid(xs, xs)
    pts = np.ascontiguousarray(np.stack([gx.ravel(), np.full(gx.size, -14.0), gz.ravel()], axis=1))
    n_ray, n_pt = len(g), len(pts)
    dev = eng.upload(g)
    fm = eng.field_prepare(dev, cfg["wavelengths"])
    d_pts = torch.from_numpy(pts).cuda()
    d_out = torch.zeros((n_pt, 6), dtype=torch.float64, device="cuda")
    pinned_pts = torch.from_numpy(pts).pin_memory()
    torch.cuda.synchronize()  # d_pts / d_out were produced on torch's stream, the library runs on its own
    for _ in range(max(args.warmup, 3)):
        fm.evaluate_device(d_pts.data_ptr(), n_pt, d_out.data_ptr())
    sampler = ClockSampler(0)
    sampler.start()
    torch.cuda.synchronize()
    ms = 0.0
    for _ in range(args.steps):
        d_out.zero_()
        torch.cuda.current_stream().synchronize()
        fm.evaluate_device(d_pts.data_ptr(), n_pt, d_out.data_ptr())
        ms += fm.last_ms
    torch.cuda.synchronize()
    sampler.stop_flag.set()
    sampler.join()
    pairs = float(n_ray) * n_pt
    value = pairs * args.steps / (ms * 1e-3)
    # e2e: host gausslets + host points in, host field out, mode fit included
    gp = eng.pinned_empty(len(g), g.dtype)
    gp[:] = g
    e2e_steps = max(1, min(args.steps, args.e2e_steps))
    t0 = time.perf_counter()
    for _ in range(e2e_steps):
        f2 = eng.field_prepare(gp, cfg["wavelengths"])
        E = f2.evaluate(pts)
        f2.free()
    e2e = pairs * e2e_steps / (time.perf_counter() - t0)
    # fp64 roofline: instructions counted from the SASS of k_field_sum per pair,
    # peak = measured DFMA issue rate of this GPU (profiles/microbench/fp64_latency.cu)
    fp64_per_pair = 142.0  # DFMA+DMUL+DADD+DSETP per pair in the SASS loop of k_field_sum (284 per 2-point iteration of 431)
    peak_inst = 63.6 * 148 * 1.965e9 / 1e12  # T fp64 lane-instructions/s, measured
    ach = fp64_per_pair * value / 1e12
    cpu = None
    if not args.no_cpu_baseline:
        cpu = fields_cpu_baseline(g, cfg["wavelengths"], pts, args)
    line = {"metric": "mode-point pairs/sec", "value": value, "unit": "pairs/s", "n_gpus": 1, "steps": args.steps,
            "warmup": max(args.warmup, 3), "ms_per_step": ms / args.steps, "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": {"workload": "fields: Michelson output gausslets on a %dx%d detector" % (side, side),
                       "n_rays": n_ray, "n_points": n_pt, "l2": "mode records 33 x 8 B per ray stream from L2/HBM; compute bound"},
            "e2e": {"value": e2e, "unit": "pairs/s", "h2d_bytes_per_step": int(g.nbytes + pts.nbytes),
                    "d2h_bytes_per_step": int(n_pt * 48), "steps": e2e_steps},
            "gpu_launches": args.steps, "clocks": sampler.summary(),
            "roofline": {"bound": "fp64", "kernel": "k_field_sum", "achieved": ach, "peak": peak_inst,
                         "unit": "T fp64 inst/s", "frac": ach / peak_inst, "traffic": None,
                         "peak_source": "measured DFMA issue rate (profiles/microbench/fp64_latency.cu)",
                         "fp64_inst_per_pair": fp64_per_pair},
            "cpu_baseline": cpu}
    print(json.dumps(line))
    return 0


def _fields_cpu_worker(a):
    flavour, g_bytes, wl, pts, t = a
    from oracle import oracle as O
    from raypier_optics_b200 import _abi as A
    core = (O.import_reference("timing") or O.import_reference("parity")) if flavour else None
    g = np.frombuffer(g_bytes, dtype=A.gausslet_dtype)
    if core is not None:
        from raypier.core import cfields
        x, y, dx, dy = O.evaluate_neighbours_gc(g)
        modes = cfields.evaluate_modes(x, y, dx, dy, blending=1.0)
        rc = O.reference_collection(core, np.ascontiguousarray(g['base_ray']), wl)
        t0 = time.perf_counter()
        cfields.sum_gaussian_modes(rc, modes, np.asarray(wl), pts, t)
        return time.perf_counter() - t0
    t0 = time.perf_counter()
    O.eval_Efield_from_gausslets(g, pts, wl)
    return time.perf_counter() - t0


def fields_cpu_baseline(g, wl, pts, args):
    """The reference's cfields.sum_gaussian_modes (OpenMP prange over points inside, :98) on all
    host cores, on a bounded sample of the same rays and points."""
    kind = cpu_kind()
    cores = os.cpu_count() or 1
    n_ray = min(len(g), 2000)
    n_pt = min(len(pts), 20000)
    os.environ.setdefault("OMP_NUM_THREADS", str(cores))
    dt = _fields_cpu_worker(("timing" if kind == "reference" else None, g[:n_ray].tobytes(), np.asarray(wl),
                             np.ascontiguousarray(pts[:n_pt]), 0.0))
    return {"value": n_ray * n_pt / dt, "unit": "pairs/s", "cores": cores, "kind": kind,
            "sample": "%d rays x %d points, reference cfields.sum_gaussian_modes with OpenMP" % (n_ray, n_pt)}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default="config5", choices=sorted(WORKLOADS))
    ap.add_argument("--rays", type=int, default=0, help="source rays per GPU (default: the config's)")
    ap.add_argument("--e2e-steps", type=int, default=5)
    ap.add_argument("--chunk-rays", type=int, default=0,
                    help="source rays per chunk (0: 131072 for rpx_trace_streamed, 2^20 gausslets / 2^22 rays for rpx_trace_consume)")
    ap.add_argument("--dropin-rays", type=int, default=200000,
                    help="source rays of the timed trace_rays drop-in call (consume-mode workloads)")
    ap.add_argument("--detector-grid", type=int, default=16, help="side of the detector grid of the consume-mode e2e arm")
    ap.add_argument("--host-source-gb", type=float, default=0.0,
                    help="cap of the pinned host source of the consume-mode e2e arm (0: min(32 GB, 45%% of MemAvailable per local rank))")
    ap.add_argument("--ref-rays-per-core", type=int, default=200000)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--capture", action="store_true",
                    help="also time trace + device-side capture plane (adds `e2e_capture`)")
    ap.add_argument("--fields", action="store_true",
                    help="measure the E-field summation (sum_gaussian_modes) instead of the trace")
    ap.add_argument("--field-grid", type=int, default=512, help="detector grid side for --fields")
    args = ap.parse_args()
    if args.fields:
        return run_fields(args)
    if args.impl == "reference":
        return run_reference_arm(args)
    if WORKLOADS[args.workload].get("mode") == "consume":
        return run_consume(args)
    if not args.chunk_rays:
        args.chunk_rays = 131072
    return run_ours(args)


if __name__ == "__main__":
    sys.exit(main())
