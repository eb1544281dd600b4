#!/usr/bin/env python
"""bench.py — headline benchmark of the wavefront path-tracing hot path (BASELINE.json metric, config C2).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl reference] [--width 2560 --height 1440]

A step = one frame (1 spp) of the C2 workload: the procedural Sponza-class atrium (~262 K triangles, 16 textures of
1024^2, 32 override-emissive lamps = 1 024 light triangles), 2560x1440, depth 4 (= 3 bounces), ReSTIR direct lighting
with temporal + spatial reuse, static camera (temporal reuse active after the warm-up frames).
  value  : Mrays/s = (extend + shadow + ReSTIR-visibility rays of the timed frames, all ranks) / time, scene resident in HBM,
           timed with CUDA events on the renderer's stream, barrier + synchronize on both sides, max over ranks.
  e2e    : the same metric through the C ABI with host buffers: every step uploads the camera pose and reads the fp32 HDR
           frame back into pinned host memory inside the timed region.
  N > 1  : sample sharding — every rank renders its own K frames with a disjoint frameCount stream into its
           fp32 accumulation buffer; ONE NCCL reduce of that buffer (inside the timed region) produces the image. scaling=weak.
  --impl reference : the CPU implementation of the same path (the oracle port; the reference itself has no CPU renderer and
           its OptiX trace cannot be built here) on all host threads, on the SAME configuration: every step one full 2560x1440 frame
           (about 3 s on 16 threads; --sample-width / --sample-height shrink it on a small host).
"""
from __future__ import annotations

import argparse
import ctypes
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

import numpy as np

METRIC = "Mrays/sec @1440p 1spp 3-bounce ReSTIR (ms/frame in ms_per_step)"
HISTORY_FRAMES = 8          # frames rendered before the warm-up on both arms: the temporal ReSTIR history of a static camera is filled
# algorithmic bytes per unit (DESIGN.md "roofline"): what a stage must move per ray / pixel, fp32 payloads, every field once
ALG_BYTES = {"raygen": 40.0, "extend": 40.0, "shadow": 44.0, "extract": 232.0, "motion": 20.0, "nee": 224.0, "bounce": 216.0,
             "ris": 256.0, "vis_gen": 288.0, "vis_trace": 32.0, "res_shade": 96.0, "temporal": 424.0, "spatial": 336.0, "combine": 368.0, "merge": 84.0}
# temporal / combine: SURVEY 8d counts the reference's 176-byte AoS SurfaceData and 80-byte Reservoir records (612 / 416 B per pixel); the SoA planes
# these kernels HAVE to move are fewer, and with the AoS figure the fraction came out above 1 (1.05 for k_temporal once it got faster). The figures used
# are the planes each pixel must read and write once: temporal = motion 8 + previous and current similarity plane 2 x 16 + the 7 other shading planes
# 112 + previous and current reservoir 2 x 80 + DIRECT channel read-modify-write 32 + reservoir store 80 = 424 (ncu: 418 B / pixel of DRAM traffic);
# combine = two reservoirs 160 + 8 shading planes 128 + store 80 = 368 (ncu: 362).


# what bounds each stage (ncu, profiles/r02_s_frame.md): goes into the `roofline` object of the stage where most of the frame goes
STAGE_NOTES = {
    "restir_ris": "instruction-issue bound (66 % issue utilisation at 16 warps / SM, 30.5 of 32 lanes active after the regrouping by survivor count): 32 candidates, 15.8 Disney BSDF evaluations per pixel on light records staged in shared memory",
    "restir_visibility": "instruction-issue bound (78 % issue utilisation, 20 of 32 lanes active): traversal of incoherent any-hit rays on an L2-resident BVH; DRAM traffic is a fifth of the algorithmic bytes because only the planes a pixel needs are touched",
    "restir_spatial": "L2 gather latency (3.5 stalled warps per issue on long scoreboard) + BSDF re-evaluation",
    "extend": "instruction-issue bound (77 %): BVH8 traversal, BVH entirely L2-resident",
    "shade": "dependent texture gathers and BSDF arithmetic at 2 blocks per SM",
}


def stage_table(stage_ms, fc, npix, depth, peak, traffic):
    """Per-stage achieved GB/s = algorithmic bytes of the stage (SURVEY 8d per-unit figure x units of this frame) / device time of the stage."""
    ext, sh, vis = fc["extend_rays"], fc["shadow_rays"], fc["visibility_rays"]
    A = ALG_BYTES
    units = {
        "raygen": ("k_raygen", npix * A["raygen"], f"{npix} rays x 40 B"),
        "extend": ("k_extend", ext * A["extend"], f"{ext} rays x 40 B (BVH traffic excluded)"),
        "shade": ("k_shade", ext * A["extract"] + npix * A["motion"] + (ext - npix) * A["nee"] + ext * A["bounce"],
                  f"{ext} hits x 232 B + {npix} px x 20 B + {ext - npix} NEE x 224 B + {ext} bounce x 216 B"),
        "shadow": ("k_shadow", sh * A["shadow"], f"{sh} rays x 44 B"),
        "restir_ris": ("k_ris", npix * A["ris"], f"{npix} px x 256 B"),
        "restir_visibility": ("k_visibility_shade", 2 * npix * (A["vis_gen"] + A["res_shade"]) + vis * A["vis_trace"], f"2 x {npix} px x (288 + 96) B + {vis} rays x 32 B"),
        "restir_temporal": ("k_temporal", npix * A["temporal"], f"{npix} px x 424 B (SoA planes; the reference's AoS records: 612 B)"),
        "restir_spatial": ("k_spatial", 2 * npix * A["spatial"], f"2 x {npix} px x 336 B"),
        "restir_combine": ("k_combine", npix * A["combine"], f"{npix} px x 368 B (SoA planes; the reference's AoS records: 416 B)"),
        "merge": ("k_merge", npix * A["merge"], f"{npix} px x 84 B"),
    }
    rows = []
    total = sum(stage_ms.values())
    for stage, ms in stage_ms.items():
        if stage not in units or ms <= 0:
            continue
        kernel, nbytes, how = units[stage]
        gbs = nbytes / (ms * 1e-3) / 1e9
        t = traffic.get(kernel)
        # dram_frac: the measured DRAM traffic of the stage (ncu capture named in traffic_capture) over this run's stage time, as a fraction of
        # the measured HBM peak — the honest utilisation figure where the algorithmic bytes (the reference's AoS records) exceed what the SoA
        # planes actually move
        rows.append({"stage": stage, "kernel": kernel, "ms_per_frame": ms, "share_of_frame": ms / total, "alg_bytes": nbytes, "alg_bytes_how": how,
                     "achieved": gbs, "unit": "GB/s", "frac": gbs / peak, "traffic": t, "dram_frac": (t / (ms * 1e-3) / 1e9 / peak) if t else None})
    return rows


def workload_scene(args):
    from lumenrenderer_b200 import scenes
    return scenes.atrium(detail=args.detail, texture_size=args.texture_size)


def settings(args, width, height, rank=0, world=1, blend=False):
    import lumenrenderer_b200 as lr
    return lr.Settings(width=width, height=height, depth=args.depth, restir=True, restir_temporal=True, restir_spatial=True,
                       blend_output=blend, device=0, first_frame_count=2 * rank, frame_count_stride=2 * world)


def rays_of(counters):
    return counters["extend_rays"] + counters["shadow_rays"] + counters["visibility_rays"]


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled DURING the timed region (B200_PROFILING.md recipe)."""
    Q = "clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap"

    def __init__(self, index):
        self.rows, self.proc = [], None
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--id={index}", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-lms", "100"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.thread = threading.Thread(target=self._pump, daemon=True); self.thread.start()
        except Exception:
            self.proc = None

    def _pump(self):
        for line in self.proc.stdout:
            self.rows.append((time.time(), line.strip()))

    def mark(self):
        return time.time()

    def stop(self, t0, t1):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        sm, mx, reasons = [], None, set()
        for t, line in self.rows:
            if t < t0 - 0.05 or t > t1 + 0.15:
                continue
            f = [x.strip() for x in line.split(",")]
            try:
                sm.append(float(f[0])); mx = float(f[1])
            except Exception:
                continue
            for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), f[3:7]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": mx, "samples": len(sm), "reasons": sorted(reasons)}


class DevPtr:
    """Exposes a raw device pointer to torch through __cuda_array_interface__ (for the NCCL reduce of the accumulation buffer)."""

    def __init__(self, ptr, nfloats):
        self.__cuda_array_interface__ = {"shape": (nfloats,), "typestr": "<f4", "data": (ptr, False), "version": 2, "strides": None}


def cpu_run(args, steps, warmup, width, height, threads=None, history=HISTORY_FRAMES):
    """The CPU port of the path (oracle) on the workload's scene at width x height: `history` + `warmup` untimed frames, then `steps` timed."""
    import __graft_entry__ as ge
    from lumenrenderer_b200 import api
    ob = ge.oracle_bindings()
    lib = ob.lib
    lib.lo_num_threads.restype = ctypes.c_int
    lib.lo_set_num_threads(int(threads or os.cpu_count() or 1))       # torchrun exports OMP_NUM_THREADS=1: ask for all host threads explicitly
    cores = int(lib.lo_num_threads())
    scene = workload_scene(args)
    r = api.Renderer(ob, settings(args, width, height))
    r.load_scene(scene)
    t_build = time.time(); r.read_lights(); t_build = time.time() - t_build       # commits the scene (BVH build, light list)
    r.render_frames(history + max(warmup, 1))
    rays, t0 = 0, time.time()
    for _ in range(steps):
        r.render_frames(1); rays += rays_of(r.frame_counters())
    dt = time.time() - t0
    r.close()
    full = (width, height) == (args.width, args.height)
    return {"mrays": rays / dt / 1e6, "ms_per_step": dt / steps * 1e3, "cores": cores, "rays_per_step": rays / steps, "scene_build_s": t_build,
            "sample": (f"the full workload: same atrium scene, {width}x{height}" if full else
                       f"same atrium scene, {width}x{height} ({width * height / (args.width * args.height):.4f} of the pixels)") +
                      f", depth {args.depth}, ReSTIR, {steps} timed frames after {history} history + {max(warmup, 1)} warm-up frames, {cores} thread(s)"}


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    w, h = args.sample_width or args.width, args.sample_height or args.height
    # history frames cost ~3 s each on the CPU: two are enough for the temporal pass to find a previous frame; the timed frames are steady state
    res = cpu_run(args, args.steps, args.warmup, w, h, history=2)
    line = {"impl": "reference", "metric": METRIC, "value": res["mrays"], "unit": "Mrays/s", "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": res["ms_per_step"], "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": config_of(args, w, h),
            "cpu_baseline": {"value": res["mrays"], "unit": "Mrays/s", "cores": res["cores"], "kind": "port", "sample": res["sample"]},
            "e2e": {"value": res["mrays"], "unit": "Mrays/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "threads": res["cores"], "scene_build_s": res["scene_build_s"],
            "note": "the reference has no CPU renderer and its ray tracing is closed-source OptiX (SURVEY 8c): this arm is the scalar C++ port of the same wavefront algorithm (oracle/), OpenMP over all host threads"}
    print(json.dumps(line), flush=True)


def config_of(args, w, h):
    """The workload, identical on both arms (the GPU arm and --impl reference): how each arm parallelises is not part of it."""
    return {"workload": f"C2 procedural Sponza-class atrium (detail {args.detail}), {w}x{h}, 1 spp, depth {args.depth} (3 bounces), ReSTIR 32 candidates + temporal + 2x spatial, static camera",
            "resolution": [w, h], "depth": args.depth, "restir": True, "textures": f"16 x {args.texture_size}^2 RGBA8",
            "l2": "per-frame working set (~0.9 KB/pixel of SoA planes, 3.3 GB at 1440p) is far larger than the 126 MB L2; no explicit flush"}


def run_gpu(args):
    import torch
    import torch.distributed as dist
    import lumenrenderer_b200 as lr

    rank, world, local = int(os.environ.get("RANK", "0")), int(os.environ.get("WORLD_SIZE", "1")), int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    W, H = args.width, args.height
    st = settings(args, W, H, rank, world, blend=world > 1)
    st.device = local
    scene = workload_scene(args)
    r = lr.Renderer(st)
    r.load_scene(scene)
    stream = torch.cuda.Stream()                    # a real (non-null) stream shared by torch events, NCCL and the renderer
    torch.cuda.set_stream(stream)
    r.set_stream(stream.cuda_stream)
    cam_pos, cam_rot = scene.camera["position"], scene.camera["rotation"]

    def barrier():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def join_comm(renderer):
        # the library's own communicator (lb_comm_*, csrc/lb_multigpu.cpp): torch.distributed only carries the 128-byte id to the other ranks
        box = [lr.comm_unique_id() if rank == 0 else None]
        dist.broadcast_object_list(box, src=0)
        renderer.comm_init(box[0], rank, world)

    if world > 1:
        warm = torch.zeros(1024, device="cuda"); dist.all_reduce(warm)          # torch's communicator (timing reductions, barrier) warmed up outside the timed region
        join_comm(r)

    r.render_frames(HISTORY_FRAMES)                 # temporal ReSTIR history of the static camera (>= 8 frames), before the warm-up proper
    r.render_frames(max(args.warmup, 3))            # >= 3 warm-up frames
    if world > 1:                                   # first use of the library's communicator outside the timed region, then a clean accumulation buffer
        r.comm_reduce_accum(0, r.accum_buffer()[2] * world); r.set_blend_mode(True)
    r.synchronize()
    counters = r.frame_counters()
    tris, lights, bvh_bytes, launches = counters["triangles"], counters["lights"], counters["bvh_bytes"], counters["kernel_launches"]
    bvh_build_ms = counters.get("bvh_build_us", 0) / 1e3

    # ---- timed region 1: device-resident throughput
    sampler = ClockSampler(local) if rank == 0 else None
    barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    t_wall0 = time.time()
    e0.record(stream)
    rays = 0
    for _ in range(args.steps):
        r.render_frames(1)
    if world > 1:
        r.comm_reduce_accum(0, args.steps * world)                              # the one collective of the path: ncclReduce inside the library + resolve on rank 0
        r.reduce_wait()                                                         # it runs on a side stream: the timed stream waits for it
    e1.record(stream)
    barrier()
    t_wall1 = time.time()
    ms = e0.elapsed_time(e1)
    rays_per_frame = rays_of(r.frame_counters())                                # static camera: every timed frame traces the same number of rays
    rays = rays_per_frame * args.steps
    clocks = sampler.stop(t_wall0, t_wall1) if sampler else None
    reduce_ms = None
    if world > 1:
        # the collective alone (59 MB of fp32 per rank at 1440p), timed on the stream after a barrier
        r.set_blend_mode(True); r.render_frames(1); barrier()
        c0, c1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        c0.record(stream); r.comm_reduce_accum(0, world); r.reduce_wait(); c1.record(stream); barrier()
        reduce_ms = c0.elapsed_time(c1)
        t = torch.tensor([ms, float(rays)], device="cuda", dtype=torch.float64)
        tmax = t.clone(); dist.all_reduce(tmax, op=dist.ReduceOp.MAX)
        tsum = t.clone(); dist.all_reduce(tsum, op=dist.ReduceOp.SUM)
        ms, rays = float(tmax[0]), float(tsum[1])

    # ---- timed region 2: end to end through the C ABI with host buffers. One GPU: camera in, HDR frame out, every step; the frame is read back
    # with the library's asynchronous read-back into two alternating pinned buffers (the copy of frame k overlaps frame k+1; every frame has
    # arrived in host memory before the clock stops). N GPUs: a step is one image of N samples — every rank uploads the camera and renders one
    # frame into a cleared accumulation buffer, the library reduces the N buffers onto rank 0 (NVLink), and ONLY rank 0 reads the image back
    # (round 1 read every rank's own frame back: 8 x 59 MB per step into one host). The reduce runs on the library's side stream and the read-back on
    # its copy stream: both overlap the next step's frame, and every image has arrived in host memory before the clock stops.
    hosts = [torch.empty((H, W, 4), dtype=torch.float32, pin_memory=True) for _ in range(2)] if rank == 0 else []
    host = hosts[0] if rank == 0 else None

    def e2e_step(k):
        if world > 1:
            r.set_blend_mode(True)                          # clears the accumulation buffer: this step's image is this step's N samples
        r.set_camera(cam_pos, cam_rot)
        r.render_frames(1)
        if world > 1:
            r.comm_reduce_accum(0, world)
        if rank == 0:
            r.readback_wait()                               # image k-1 is now in hosts[(k-1) % 2]
            r.read_hdr_async(hosts[k % 2].data_ptr(), hosts[0].numel() * 4)

    e2e_step(0)
    if rank == 0:
        r.readback_wait()
    barrier()
    t0 = time.time()
    for k in range(args.steps):
        e2e_step(k)
    if rank == 0:
        r.readback_wait()
    torch.cuda.synchronize()
    e2e_s = time.time() - t0
    host = hosts[(args.steps - 1) % 2] if rank == 0 else None
    e2e_rays = rays_per_frame * args.steps
    if world > 1:
        t = torch.tensor([e2e_s], device="cuda", dtype=torch.float64); dist.all_reduce(t, op=dist.ReduceOp.MAX); e2e_s = float(t[0])
        t = torch.tensor([float(e2e_rays)], device="cuda", dtype=torch.float64); dist.all_reduce(t, op=dist.ReduceOp.SUM); e2e_rays = float(t[0])
        r.set_blend_mode(False)
    finite = bool(np.isfinite(host.numpy()).all()) if rank == 0 else True

    # ---- per-stage device time (CUDA events on the renderer's stream, per frame) for the roofline lines
    # Stage times are taken with the two chains of a frame serialised (lb_set_overlap(0), the default): every stage time is an exclusive
    # device time. LB_OVERLAP=1 runs the timed regions above with the ReSTIR chain and the bounce chain on two streams (an experiment).
    r.set_overlap(0)
    r.render_frames(1)
    stage_ms, stage_frames = {}, max(3, min(args.steps, 10))
    for _ in range(stage_frames):
        r.render_frames(1)
        for k, v in r.frame_stats().items():
            stage_ms[k] = stage_ms.get(k, 0.0) + v / 1e3
    stage_ms = {k: v / stage_frames for k, v in stage_ms.items()}
    serial_ms = sum(stage_ms.values())
    fc = r.frame_counters()

    extras = None if args.no_extras else run_extras(args, lr, torch, dist, rank, world, local, scene, stream, join_comm, barrier)

    if rank == 0:
        peaks = {}
        try:
            peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
        except Exception:
            pass
        peak = float(peaks.get("hbm_gbs", 6650.0))
        traffic, traffic_capture = {}, None
        try:    # dram__bytes_read.sum + dram__bytes_write.sum per frame and kernel, from the committed `ncu --set full` capture of this bench command
            tj = json.load(open(os.path.join(ROOT, "profiles", "traffic.json")))
            traffic, traffic_capture = tj["dram_bytes_per_frame"], {"source": tj.get("source"), "note": tj.get("note")}
        except Exception:
            pass
        stage_ms["restir"] = sum(v for k, v in stage_ms.items() if k.startswith("restir_"))
        detail = {k: v for k, v in stage_ms.items() if k != "restir"}
        table = stage_table(detail, fc, W * H, args.depth, peak, traffic)
        by_stage = {r["stage"]: r for r in table}
        top = max(table, key=lambda r: r["ms_per_frame"])
        ext = by_stage["extend"]
        peak_source = "MEASURED_PEAKS.json hbm_gbs" if peaks else "fallback 6650 GB/s (B200_PROFILING.md)"
        # `roofline`: the kernel where most of the frame goes. `roofline_extend`: the traversal kernel the north star asks to be reported
        # against HBM (it is latency / issue bound on an L2-resident BVH; see profiles/ for the stall breakdown).
        roofline = {"kernel": top["kernel"], "stage": top["stage"], "bound": "hbm", "achieved": top["achieved"], "peak": peak, "unit": "GB/s", "frac": top["frac"],
                    "peak_source": peak_source, "traffic": top["traffic"], "traffic_capture": traffic_capture, "dram_frac": top["dram_frac"], "alg_bytes_per_frame": top["alg_bytes"], "alg_bytes_how": top["alg_bytes_how"],
                    "ms_per_frame": top["ms_per_frame"], "share_of_frame": top["share_of_frame"],
                    "note": STAGE_NOTES.get(top["stage"], "") + "; reported against HBM because no stage is a dense contraction"}
        roofline_extend = {"kernel": "k_extend (BVH8 traversal, all waves of a frame)", "bound": "hbm", "achieved": ext["achieved"], "peak": peak, "unit": "GB/s", "frac": ext["frac"],
                           "peak_source": peak_source, "traffic": ext["traffic"], "traffic_capture": traffic_capture, "dram_frac": ext["dram_frac"], "alg_bytes_per_ray": ALG_BYTES["extend"], "rays_per_launch_avg": fc["extend_rays"] / args.depth,
                           "ms_per_frame": ext["ms_per_frame"], "share_of_frame": ext["share_of_frame"], "mrays_per_s": fc["extend_rays"] / (ext["ms_per_frame"] * 1e-3) / 1e6}
        # the whole frame against HBM: every stage's algorithmic bytes over the device time of a frame (the timed region above, not the sum of stages)
        frame_ms = ms / args.steps
        alg_total = sum(r["alg_bytes"] for r in table)
        dram_total = sum(r["traffic"] for r in table if r["traffic"]) or None
        roofline_frame = {"bound": "hbm", "alg_bytes_per_frame": alg_total, "achieved": alg_total / (frame_ms * 1e-3) / 1e9, "peak": peak, "unit": "GB/s",
                          "frac": alg_total / (frame_ms * 1e-3) / 1e9 / peak, "traffic": dram_total, "traffic_capture": traffic_capture,
                          "dram_frac": (dram_total / (frame_ms * 1e-3) / 1e9 / peak) if dram_total else None, "ms_per_frame": frame_ms}
        cpu = None
        if world == 1 and not args.no_cpu_baseline:
            # all host threads on the full workload (3 frames of about 3 s), one thread on 1/16 of the pixels (3 frames of about 2.5 s)
            c = cpu_run(args, 3, 1, W, H, history=2)
            c1 = cpu_run(args, 3, 1, W // 4, H // 4, threads=1, history=2)
            cpu = {"value": c["mrays"], "unit": "Mrays/s", "cores": c["cores"], "kind": "port", "sample": c["sample"], "ms_per_sample_frame": c["ms_per_step"],
                   "single_thread": {"value": c1["mrays"], "unit": "Mrays/s", "cores": 1, "sample": c1["sample"], "ms_per_sample_frame": c1["ms_per_step"]}}
        line = {"metric": METRIC, "value": rays / (ms * 1e-3) / 1e6, "unit": "Mrays/s", "n_gpus": world, "steps": args.steps, "warmup": max(args.warmup, 3),
                "ms_per_step": ms / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
                "config": config_of(args, W, H), "parallelism": f"sample-sharded x{world}", "history_fill_frames": HISTORY_FRAMES,
                "fps": args.steps / (ms * 1e-3), "samples_per_s": W * H * args.steps * world / (ms * 1e-3), "rays_per_frame": rays_per_frame,
                "scene": {"triangles": tris, "lights": lights, "bvh_bytes": bvh_bytes, "bvh_build_ms": bvh_build_ms, "bvh_first_alloc_ms": counters.get("bvh_alloc_us", 0) / 1e3, "bvh_builder": os.environ.get("LB_BVH_BUILDER", "ploc")},
                "e2e": {"value": e2e_rays / e2e_s / 1e6, "unit": "Mrays/s", "h2d_bytes_per_step": 28 * world, "d2h_bytes_per_step": W * H * 16, "ms_per_step": e2e_s / args.steps * 1e3,
                        "what": "one GPU: camera upload + frame + HDR read-back per step" if world == 1 else
                                f"per step: camera upload and one frame on each of the {world} ranks, ncclReduce of the accumulation buffers onto rank 0 inside the library, HDR read-back on rank 0 only"},
                "reduce_ms": reduce_ms, "extras": extras,
                "gpu_launches": launches * args.steps, "launches_per_frame": launches,
                "overlap": {"mode": int(os.environ.get("LB_OVERLAP", "5")), "ms_per_frame_serialised": serial_ms,
                            "note": "mask: bit 0 = shadow rays of bounce wave d on a side stream under the extend launch of wave d+1; bit 2 = ReSTIR chain launched after the first bounce wave, later waves beside it (default 5); bit 1 (off, measured slower) = ReSTIR chain beside all bounce waves; stage_ms / roofline_kernels are always exclusive times measured with mode 0"},
                "roofline": roofline, "roofline_frame": roofline_frame, "roofline_extend": roofline_extend, "roofline_kernels": table, "stage_ms": stage_ms, "cpu_baseline": cpu, "clocks": clocks, "output_finite": finite}
        print(json.dumps(line), flush=True)
    r.close()
    if world > 1:
        dist.destroy_process_group()


def scenes_quat_y(deg):
    from lumenrenderer_b200 import scenes
    return scenes._quat_y(deg)


def run_extras(args, lr, torch, dist, rank, world, local, scene, stream, join_comm, barrier):
    """Two more measurements on the same ranks, reported beside the headline (never instead of it):
    strong_bands_ms_per_frame — ONE 2560x1440 frame per step, split into row bands with a 60-row ReSTIR halo across the ranks and gathered on
        rank 0 by the library (lb_band_settings + lb_comm_gather_bands): strong scaling / single-frame latency;
    c5_time_to_64spp_s — BASELINE configs[4]: 3840x2160 progressive, 64 samples per pixel in total, sample-sharded (64 / N frames per rank) with
        one reduce of the fp32 accumulation buffer at the end."""
    out = {}
    W, H = args.width, args.height
    def timed(fn, steps):
        barrier()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(stream); fn(steps); b.record(stream); barrier()
        ms = a.elapsed_time(b)
        if world > 1:
            t = torch.tensor([ms], device="cuda", dtype=torch.float64); dist.all_reduce(t, op=dist.ReduceOp.MAX); ms = float(t[0])
        return ms
    # ---- row bands
    full_st = settings(args, W, H)
    st, (y0, y1) = lr.band_settings(full_st, rank, world)
    st.device = local
    rb = lr.Renderer(st); rb.load_scene(scene); rb.set_stream(stream.cuda_stream)
    full = torch.zeros((H, W, 4), device="cuda") if rank == 0 else None
    if world > 1:
        join_comm(rb)
    def band_frames(n):
        for _ in range(n):
            rb.render_frames(1)
            if world > 1:
                rb.comm_gather_bands(0, full.data_ptr() if rank == 0 else 0)
    band_frames(HISTORY_FRAMES + 3)
    steps = max(3, min(args.steps, 10))
    ms = timed(band_frames, steps)
    rows = torch.tensor([float(st.height)], device="cuda", dtype=torch.float64)
    if world > 1:
        dist.all_reduce(rows, op=dist.ReduceOp.SUM)
    out["strong_bands_ms_per_frame"] = ms / steps
    out["strong_bands"] = {"frames": steps, "rows_rendered_over_rows_owned": float(rows[0]) / H, "halo_rows": 60, "gather_bytes_per_frame": (H - (y1 - y0)) * W * 16 if rank == 0 else None,
                           "finite": bool(torch.isfinite(full).all()) if (rank == 0 and world > 1) else True}
    if world > 1:
        rb.synchronize(); rb.comm_destroy()
    rb.close(); del full
    # ---- the same workload seen along the atrium's long axis. The headline camera (unchanged since round 1 so that rounds stay comparable) stands
    # 2.5 m in front of the atrium's end wall and faces it: 92 % of its primary hits are that one clear-coated wall. This view looks the other way,
    # down the colonnade (floor, columns, arches, drapes, all 24 materials): the more Sponza-like picture, reported beside the headline.
    if rank == 0:
        ra = lr.Renderer(settings(args, W, H)); ra.load_scene(scene); ra.set_stream(stream.cuda_stream)
        ra.set_camera(scene.camera["position"], scenes_quat_y(80.0))
        ra.render_frames(HISTORY_FRAMES + 3); ra.synchronize()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        n = max(3, min(args.steps, 10))
        a.record(stream); ra.render_frames(n); b.record(stream); torch.cuda.synchronize()
        fc = ra.frame_counters()
        ra.set_overlap(0); ra.render_frames(2)
        out["alt_view"] = {"camera": "same position, yaw +80 deg (looking down the colonnade)", "ms_per_frame": a.elapsed_time(b) / n, "rays_per_frame": rays_of(fc),
                           "mrays_per_s": rays_of(fc) / (a.elapsed_time(b) / n) / 1e3, "stage_ms": {k: v / 1e3 for k, v in ra.frame_stats().items()}}
        ra.close()
    # ---- C5: 4K, 64 spp in total
    total_spp = 64
    frames = total_spp // world + (1 if rank < total_spp % world else 0)
    st5 = lr.shard_settings(settings(args, 3840, 2160), rank, world); st5.device = local
    r5 = lr.Renderer(st5); r5.load_scene(scene); r5.set_stream(stream.cuda_stream)
    if world > 1:
        join_comm(r5)
    r5.render_frames(2)
    if world > 1:
        r5.comm_reduce_accum(0, 2 * world)
    r5.set_blend_mode(True)                                 # warm-up done: a clean accumulation buffer
    def c5(_):
        r5.render_frames(frames)
        if world > 1:
            r5.comm_reduce_accum(0, total_spp); r5.reduce_wait()
        else:
            r5.resolve_accum(total_spp)
    ms5 = timed(c5, 1)
    out["c5_time_to_64spp_s"] = ms5 / 1e3
    out["c5"] = {"resolution": [3840, 2160], "total_spp": total_spp, "frames_on_this_rank": frames, "reduce_bytes": 3840 * 2160 * 16 if world > 1 else 0}
    if world > 1:
        r5.synchronize(); r5.comm_destroy()
    r5.close()
    # ---- the reference's own Sponza asset, when it is on this box (it is not part of the repository: LB_SPONZA_GLTF or tmp_assets/)
    if world == 1:
        path = os.environ.get("LB_SPONZA_GLTF") or os.path.join(ROOT, "tmp_assets", "Sponza", "Sponza.gltf")
        if os.path.exists(path):
            from lumenrenderer_b200 import scenes
            sp_scene, cam_pos, cam_rot, info, t_load = scenes.real_sponza(path)
            rs = lr.Renderer(settings(args, W, H)); rs.load_scene(sp_scene); rs.set_stream(stream.cuda_stream); rs.set_camera(cam_pos, cam_rot)
            rs.render_frames(HISTORY_FRAMES + 3)
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            n = max(3, min(args.steps, 10))
            a.record(stream); rs.render_frames(n); b.record(stream); torch.cuda.synchronize()
            fc = rs.frame_counters()
            rs.set_overlap(0); rs.render_frames(2); rs.synchronize()
            rays = fc["extend_rays"] + fc["shadow_rays"] + fc["visibility_rays"]
            ms = a.elapsed_time(b) / n
            out["sponza_real"] = {"asset": os.path.basename(path), "triangles": info.get("triangles"), "images": info.get("images"), "undecoded_images": info.get("undecoded_images"),
                                  "load_and_decode_s": t_load, "ms_per_frame": ms, "mrays_per_s": rays / ms / 1e3, "rays_per_frame": rays,
                                  "bvh_build_ms": fc["bvh_build_us"] / 1e3, "stage_ms": {k: v / 1e3 for k, v in rs.frame_stats().items()},
                                  "parity": "profiles/sponza_real.py compares this scene with the oracle at 480x270 (profiles/r02_sponza_real.json)"}
            rs.close()
        else:
            out["sponza_real"] = {"skipped": "Sponza.gltf is not on this box (set LB_SPONZA_GLTF); the run with the asset is recorded in profiles/r02_sponza_real.json"}
    return out


def run_bands(args):
    """--mode bands: ONE frame split into row bands across the ranks (single-frame latency; strong scaling). Every rank
    renders its rows + a 60-row ReSTIR halo, then the owned rows are gathered on rank 0 device to device (NCCL point-to-point on
    the renderer's stream) inside the timed region. A step = one complete 2560x1440 frame assembled on rank 0."""
    import torch
    import torch.distributed as dist
    import lumenrenderer_b200 as lr
    from lumenrenderer_b200 import sharding

    rank, world, local = int(os.environ.get("RANK", "0")), int(os.environ.get("WORLD_SIZE", "1")), int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    W, H = args.width, args.height
    st, (y0, y1, h0, h1) = sharding.band_settings(settings(args, W, H), rank, world)
    st.device = local
    bands = sharding.band_partition(H, world, width=W)
    scene = workload_scene(args)
    r = lr.Renderer(st)
    r.load_scene(scene)
    stream = torch.cuda.Stream(); torch.cuda.set_stream(stream); r.set_stream(stream.cuda_stream)
    ptr, nbytes = r.hdr_buffer()
    rows = torch.as_tensor(DevPtr(ptr, nbytes // 4), device=torch.device("cuda", local)).view(h1 - h0, W, 4)
    full = torch.zeros((H, W, 4), device="cuda") if rank == 0 else None

    def barrier():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    for _ in range(max(args.warmup, 3)):
        r.render_frames(1); sharding.gather_bands(rows, full, bands, rank)
    sampler = ClockSampler(local) if rank == 0 else None
    barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    t0 = time.time(); e0.record(stream)
    for _ in range(args.steps):
        r.render_frames(1)
        sharding.gather_bands(rows, full, bands, rank)
    e1.record(stream)
    barrier(); t1 = time.time()
    ms = e0.elapsed_time(e1)
    fc = r.frame_counters()
    rays = float(rays_of(fc)) * args.steps
    clocks = sampler.stop(t0, t1) if sampler else None
    if world > 1:
        t = torch.tensor([ms], device="cuda", dtype=torch.float64); dist.all_reduce(t, op=dist.ReduceOp.MAX); ms = float(t[0])
        t = torch.tensor([rays], device="cuda", dtype=torch.float64); dist.all_reduce(t, op=dist.ReduceOp.SUM); rays = float(t[0])
    if rank == 0:
        finite = bool(torch.isfinite(full).all()) and float(full[..., :3].sum()) > 0
        rendered_rows = sum(b[3] - b[2] for b in bands)
        line = {"metric": METRIC, "mode": "bands", "value": rays / (ms * 1e-3) / 1e6, "unit": "Mrays/s", "n_gpus": world, "steps": args.steps, "warmup": max(args.warmup, 3),
                "ms_per_step": ms / args.steps, "fps": args.steps / (ms * 1e-3), "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
                "config": dict(config_of(args, W, H), parallelism=f"row bands x{world} with a {sharding.RESTIR_HALO}-row ReSTIR halo, gather on rank 0"),
                "bands": [list(b) for b in bands], "rows_rendered_over_rows_owned": rendered_rows / H, "gather_bytes_per_step": (H - (y1 - y0)) * W * 16,
                "gpu_launches": fc["kernel_launches"] * args.steps, "clocks": clocks, "output_finite": finite}
        print(json.dumps(line), flush=True)
    r.close()
    if world > 1:
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=8)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--width", type=int, default=2560)
    ap.add_argument("--height", type=int, default=1440)
    ap.add_argument("--depth", type=int, default=4)
    ap.add_argument("--detail", type=float, default=0.78)
    ap.add_argument("--texture-size", type=int, default=1024)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-extras", action="store_true", help="skip the extra measurements (row bands of one frame across the ranks, C5 time to 64 spp at 4K)")
    ap.add_argument("--sample-width", type=int, default=0, help="--impl reference: render this width instead of --width (a bounded sample for small hosts)")
    ap.add_argument("--sample-height", type=int, default=0)
    ap.add_argument("--mode", default="samples", choices=["samples", "bands"],
                    help="multi-GPU partitioning: independent sample streams + one reduce (default, weak scaling) or row bands of one frame + one gather (strong scaling)")
    args = ap.parse_args()
    if args.impl == "reference":
        run_reference(args)
    elif args.mode == "bands":
        run_bands(args)
    else:
        run_gpu(args)


if __name__ == "__main__":
    main()
