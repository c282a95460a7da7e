#!/usr/bin/env python
"""bench.py -- headline benchmark of the B200-native CudaRaster hot path.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl b200|reference] [--workload c2|c3|c4]

A "step" is one pass of the hot path (CudaRaster::drawTriangles: setup -> bin -> coarse -> fine,
reference src/cudaraster/CudaRaster.cpp:237-342) over one synthetic frame.  Default workload is
BASELINE.json configs[1] ("c2"): 1 M-triangle tessellated grid, Gouraud, depth test, 1920x1080.

N = 1   the frame is rendered K times on cuda:0.
N > 1   one process per GPU (torchrun); view-parallel sharding: every step each rank
        renders one frame of the workload (the same frame on every rank, so that per-GPU work is exactly
        the N = 1 workload) and the colour frames are composited on rank 0 (peer memory over NVLink, or
        NCCL; the composite is inside the timed region).  Weak scaling.  The 48-view round-robin batch and
        the 4K sort-first frame of BASELINE config 5 are tools/views_48.py and tools/sort_first_4k.py.

Prints ONE JSON line (rank 0).  Keys: see the task contract; additionally
  value        N = 1: the metric SURVEY.md 8(d) defines -- T / t_frame, t_frame = the sum of the four stage intervals between CUDA
               events at the reference's five positions (CudaRaster.cpp:593-661), ONE frame in flight, median over the K frames;
               N > 1: N * T / (K-frame bracket / K), same frames (one in flight per rank, stage events on), max over ranks,
               composite inside
  value_unbroken_chain / value_two_in_flight   the same K frames without stage events (one kernel chain) / alternating
               between two contexts (double-buffered swap chain): throughput modes, not the headline
  enqueue_ms_per_step  host time spent enqueueing one frame (a number near ms_per_step means the run is host-bound)
  stage_ms     mean per-stage device time (the reference's five event positions)
  roofline     dominant kernel: algorithmic bytes per launch / mean launch duration vs measured HBM peak
  cpu_baseline the CPU oracle (oracle/golden.hpp) timed on this box's host cores on a bounded sample
  ref_kernels  the reference's own CUDA kernels rebuilt for sm_100a with the synchronisation patch (oracle/_ref/libcrref_cuda_sync.so),
               frame checked against the oracle in the same run, and the speed-up of this pipeline over them
  e2e          the same metric through crb_draw_triangles_host (pinned HOST buffers, H2D + D2H inside)

--impl reference times the CPU restatement of the path (the reference has no runnable CPU path of
its own: its host emulators are broken in this port) with all host threads.
"""
import argparse
import ctypes
import json
import os
import statistics
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

import numpy as np  # noqa: E402

WORKLOADS = {
    # name: (description, scene fn name, kwargs, width, height, shader, samplesLog2, flags, varyings)
    "c2": ("C2: 1M-triangle tessellated grid, Gouraud, depth test, 1920x1080", "grid_gouraud", {"nx": 1000, "ny": 500}, 1920, 1080, "gouraud", 0, 3, 1),
    "c3": ("C3: 5M-triangle 5-layer scene, Phong + procedural texture, 4x MSAA, 2048x2048", "layered_phong", {"nx": 1000, "ny": 500, "layers": 5}, 2048, 2048, "texPhong", 2, 3, 3),
    "c4": ("C4: 10M sub-pixel triangles, PassThrough, depth test, 1920x1080", "subpixel_soup", {"num_tris": 10_000_000}, 1920, 1080, "passthrough", 0, 1, 0),
}
STAGES = ("triangleSetup", "binRaster", "coarseRaster", "fineRaster")
NUM_INPUT_COPIES = 4  # inputs rotate over this many copies so that no step finds its inputs in L2


def measured_peak():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    try:
        return float(json.load(open(p))["hbm_gbs"]), "measured (MEASURED_PEAKS.json)"
    except Exception:
        return 6650.0, "fallback (B200_PROFILING.md)"


class ClockSampler(threading.Thread):
    """Samples SM clocks and throttle reasons with NVML while the timed region runs."""

    def __init__(self, index):
        super().__init__(daemon=True)
        self.index, self.samples, self.reasons, self.max_mhz, self.stop_flag = index, [], set(), None, False
        try:
            import pynvml
            pynvml.nvmlInit()
            self.nv = pynvml
            self.h = pynvml.nvmlDeviceGetHandleByIndex(index)
            self.max_mhz = pynvml.nvmlDeviceGetMaxClockInfo(self.h, pynvml.NVML_CLOCK_SM)
        except Exception:
            self.nv = None

    def run(self):
        if self.nv is None:
            return
        nv = self.nv
        names = {nv.nvmlClocksThrottleReasonHwSlowdown: "hw_slowdown", nv.nvmlClocksThrottleReasonHwThermalSlowdown: "hw_thermal_slowdown",
                 nv.nvmlClocksThrottleReasonSwThermalSlowdown: "sw_thermal_slowdown", nv.nvmlClocksThrottleReasonSwPowerCap: "sw_power_cap",
                 nv.nvmlClocksThrottleReasonHwPowerBrakeSlowdown: "hw_power_brake"}
        while not self.stop_flag:
            try:
                self.samples.append(nv.nvmlDeviceGetClockInfo(self.h, nv.NVML_CLOCK_SM))
                r = nv.nvmlDeviceGetCurrentClocksThrottleReasons(self.h)
                for bit, name in names.items():
                    if r & bit:
                        self.reasons.add(name)
            except Exception:
                pass
            time.sleep(0.05)

    def result(self):
        self.stop_flag = True
        if self.is_alive():
            self.join(timeout=1.0)
        if not self.samples:
            return {"sm_mhz": None, "sm_max_mhz": self.max_mhz, "reasons": ["unavailable"]}
        return {"sm_mhz": statistics.median(self.samples), "sm_max_mhz": self.max_mhz, "reasons": sorted(self.reasons)}


def make_scene(workload):
    import cudaraster_linux_b200 as crb
    desc, fn, kw, w, h, shader, s_log2, flags, k_var = WORKLOADS[workload]
    verts, idx = getattr(crb.scenes, fn)(**kw)
    return desc, verts, idx, w, h, shader, s_log2, flags, k_var


def config_dict(workload, n_gpus):
    """The `config` of the JSON line -- the same dict in both arms (b200 and reference)."""
    desc, fn, kw, w, h, shader, s_log2, flags, k_var = WORKLOADS[workload]
    import cudaraster_linux_b200 as crb
    n_tris = {"c2": 1_000_000, "c3": 5_000_000, "c4": 10_000_000}[workload]
    return {"workload": desc, "triangles": n_tris, "resolution": [w, h], "samples": 1 << s_log2, "pipe": crb.pipe_name(shader, s_log2, flags, "BlendReplace"),
            "sharding": "1 GPU" if n_gpus == 1 else "view-parallel: 1 frame of the workload per rank per step (the same frame on every rank: per-GPU work = the N = 1 workload), colour frames composited on rank 0 inside the timed region",
            "l2": "GPU arm: inputs rotate over %d device copies and each frame rewrites its intermediates (setup records, queues, surfaces): working set > 126 MB L2; "
                  "CPU arm: one frame per step" % NUM_INPUT_COPIES}


def stage_bytes(counts, num_tris, k_var, pixels, samples, lerp):
    """Per-stage split of B_alg (DESIGN.md 'algorithmic bytes'); the four add up to B_alg."""
    c = counts
    t, tsub, v = num_tris, c["numSubtris"], c["vertsReferenced"]
    return {
        "triangleSetup": 12 * t + 16 * v + 1 * t + 80 * tsub,
        "binRaster": 1 * t + 16 * tsub + 4 * c["eBin"],
        "coarseRaster": (4 + 16) * c["eBin"] + 4 * c["eTile"],
        "fineRaster": (4 + 16) * c["eTile"] + 16 * c["eCov"] + (48 * c["eShade"] if lerp else 0) + 16 * k_var * v + 8 * pixels * samples,
    }


def cpu_oracle_frame(verts, idx, w, h, shader, s_log2, flags, threads, want_counts):
    from oracle import binding as G
    from tests.util import STRIDE
    cfg = G.make_config(w, h, s_log2, flags, STRIDE[shader], shader, "BlendReplace", clear=G.clear_values(), threads=threads)
    t0 = time.perf_counter()
    r = G.render(cfg, verts, idx, want_counts=want_counts)
    return time.perf_counter() - t0, r


def run_reference_arm(args, emit):
    """The CPU restatement of the path on the host cores (rank 0 only)."""
    if int(os.environ.get("RANK", "0")) != 0:
        return 0
    from oracle import binding as G
    desc, verts, idx, w, h, shader, s_log2, flags, k_var = make_scene(args.workload)
    threads = G.hardware_threads()
    n = idx.shape[0]
    # bounded sample: ONE whole frame per step (the oracle renders C2 / C3 / C4 in 1 / 3 / 2 s on 8 threads)
    sample = n
    for _ in range(1 if args.warmup > 0 else 0):
        cpu_oracle_frame(verts, idx[:sample], w, h, shader, s_log2, flags, threads, False)
    times = [cpu_oracle_frame(verts, idx[:sample], w, h, shader, s_log2, flags, threads, False)[0] for _ in range(args.steps)]
    ms = 1e3 * statistics.median(times)
    value = sample / (ms * 1e-3) / 1e6
    line = {
        "impl": "reference", "metric": "Mtris/s", "value": value, "unit": "Mtris/s", "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": ms, "frames_per_s": 1e3 / ms * (sample / n), "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "s32/f32",
        "data": "synthetic", "config": config_dict(args.workload, args.gpus),
        "cpu_baseline": {"value": value, "unit": "Mtris/s", "cores": threads, "kind": "port",
                         "sample": "one whole frame (%d triangles) per step, median of %d, all %d host threads (oracle/golden.hpp; the reference has no runnable CPU path)" % (sample, args.steps, threads)},
        "e2e": {"value": value, "unit": "Mtris/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    emit(line)
    return 0


def time_ref_kernels(workload, timeout=240):
    """Times the reference's own CUDA kernels rebuilt for sm_100a, with the synchronisation patch that lets them run on
    Blackwell (oracle/ref_kernels/b200_sync_patch.py), in a SUBPROCESS (Fermi code: it may still hang or fault).  The frame is
    checked against the CPU oracle in the same run: a number only counts when `status` says pixel-exact."""
    lib = os.path.join(ROOT, "oracle", "_ref", "libcrref_cuda_sync.so")
    if not os.path.exists(lib):
        return {"status": "not built"}
    try:
        r = subprocess.run([sys.executable, os.path.join(ROOT, "oracle", "run_ref_kernels.py"), "--workload", workload, "--frames", "23", "--check"],
                           capture_output=True, text=True, timeout=timeout, env=dict(os.environ, CRREF_LIBRARY=lib))
    except subprocess.TimeoutExpired:
        return {"status": "hang (killed after %d s)" % timeout}
    for ln in reversed(r.stdout.strip().splitlines()):
        if ln.startswith("{"):
            d = json.loads(ln)
            if d.get("status") == "ok":
                d["status"] = "sync-patched, pixel-exact (depth and colour equal to the oracle's frame)"
            return d
    return {"status": "failed rc=%d: %s" % (r.returncode, (r.stderr or r.stdout)[-300:].replace("\n", " | "))}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=30)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--workload", default="c2", choices=sorted(WORKLOADS))
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-ref-kernels", action="store_true")
    ap.add_argument("--no-stagger", action="store_true", help="N > 1, push composite: all ranks start their frame loops in phase")
    ap.add_argument("--frames-in-flight", type=int, default=2, choices=[1, 2],
                    help="2 = consecutive frames alternate between two contexts / streams / surface sets, so the (issue-bound) setup of one frame "
                         "overlaps the (latency-bound) fine raster of the other -- a double-buffered swap chain; 1 = one context, one stream")
    ap.add_argument("--composite", default="auto", choices=["auto", "peer", "push", "nccl"],
                    help="N > 1: 'peer' = every rank renders straight into rank 0's memory over NVLink (CUDA IPC, no gather); 'push' = render locally, "
                         "DMA copy of finished frames into rank 0's slots on a side stream; 'nccl' = NCCL gather of finished frames; 'auto' = peer for 2 GPUs, push beyond "
                         "(measured: with 8 ranks the simultaneous fine-raster stores exceed rank 0's NVLink ingress, the DMA copies spread over the frame)")
    args = ap.parse_args()
    # stdout carries the ONE JSON line and nothing else: libraries that write to file descriptor 1 (NCCL's version banner,
    # nvcc of the reference-kernel harness) are pointed at stderr; emit() writes the line to the real stdout.
    sys.stdout.flush()
    real_stdout = os.dup(1)
    os.dup2(2, 1)

    def emit(line):
        os.write(real_stdout, (json.dumps(line) + "\n").encode())

    if args.impl == "reference":
        return run_reference_arm(args, emit)
    args.warmup = max(args.warmup, 3)

    import torch
    import torch.distributed as dist
    import cudaraster_linux_b200 as crb
    from cudaraster_linux_b200 import multigpu

    rank, world, local = int(os.environ.get("RANK", "0")), int(os.environ.get("WORLD_SIZE", "1")), int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device -- the product has no CPU path (use --impl reference for the CPU oracle)")
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        os.environ.setdefault("NCCL_DEBUG_FILE", "/dev/stderr")   # NCCL's version banner / debug lines must not precede the JSON line on stdout
        dist.init_process_group("nccl", device_id=dev)

    desc, verts, idx, w, h, shader, s_log2, flags, k_var = make_scene(args.workload)
    n_tris, n_samples = idx.shape[0], 1 << s_log2
    raster = crb.CudaRaster(local)
    color = crb.CudaSurface((w, h), crb.CudaSurface.FORMAT_RGBA8, n_samples, device=dev)
    depth = crb.CudaSurface((w, h), crb.CudaSurface.FORMAT_DEPTH32, n_samples, device=dev)
    raster.setSurfaces(color, depth)
    raster.setPixelPipe(None, crb.pipe_name(shader, s_log2, flags, "BlendReplace"))

    # view-parallel sharding: one frame per rank per step
    h_verts = torch.from_numpy(verts).pin_memory()
    h_idx = torch.from_numpy(idx).pin_memory()
    copies = []
    for k in range(NUM_INPUT_COPIES):
        vv = verts   # every rank renders the C2 frame itself: per-GPU work is exactly the N = 1 workload (weak scaling); the 48-view batch of config 5(ii) is tools/views_48.py
        copies.append((torch.from_numpy(vv).to(dev), torch.from_numpy(idx).to(dev)))
    # N > 1, composite to rank 0.  Default: rank 0 owns two frame slots per rank, exported with CUDA IPC; every
    # rank renders STRAIGHT into its slot, so the fine raster's colour stores cross NVLink / NVSwitch while the frame is being
    # rendered and there is no gather step at all.  --composite nccl: two local colour surfaces and an NCCL gather of frame k
    # (side stream) that overlaps the rendering of frame k+1.
    if args.composite == "auto":
        args.composite = "peer" if world <= 2 else "push"
    peer = world > 1 and args.composite == "peer"
    push = world > 1 and args.composite == "push"
    sink = None
    if peer or push:
        try:
            sink = multigpu.PeerFrameSink(world, rank, color.tensor.numel() * 4, depth=2, device=dev)
        except multigpu.PeerMemoryUnavailable as e:   # raised on every rank: fall back to the NCCL gather together
            print("bench.py: %s -- falling back to --composite nccl" % e, file=sys.stderr)
            args.composite, peer, push = "nccl", False, False
    if world > 1 and args.composite == "nccl":
        args.frames_in_flight = 1   # the gather's own side stream already overlaps frame k with frame k+1; two render lanes only get in its way (N=2: 16.2 vs 18.8 Gtris/s)
    peer_surfaces = [sink.surface(k, (w, h), n_samples) for k in range(2)] if peer else None
    # Frames in flight: consecutive frames are independent (own view / own surfaces), so they alternate between F contexts, each
    # with its own stream, work buffers and depth surface -- a double-buffered swap chain.  Triangle setup is issue-bound and the
    # fine raster latency-bound; two frames in flight fill each other's gaps (tools/probe_overlap.py: +17 % on C2).
    F = args.frames_in_flight
    stream = torch.cuda.current_stream(dev)
    rasters, lane_streams, depths = [raster], [stream], [depth]
    for _ in range(F - 1):
        r2 = crb.CudaRaster(local)
        r2.setPixelPipe(None, crb.pipe_name(shader, s_log2, flags, "BlendReplace"))
        rasters.append(r2)
        lane_streams.append(torch.cuda.Stream(device=dev))
        depths.append(crb.CudaSurface((w, h), crb.CudaSurface.FORMAT_DEPTH32, n_samples, device=dev))
    colors = [color] + ([crb.CudaSurface((w, h), crb.CudaSurface.FORMAT_RGBA8, n_samples, device=dev)] if (F > 1 or world > 1) and not peer else [])
    gatherer = multigpu.AsyncFrameGather([c.tensor for c in colors], world, rank, dst=0) if world > 1 and not (peer or push) else None
    if push:
        sink.attach_local([c.tensor for c in colors])

    def step(k, asynchronous=True, lanes=True):
        lane = k % F if lanes else 0
        r = rasters[lane]
        vb, ib = copies[k % NUM_INPUT_COPIES]
        with torch.cuda.stream(lane_streams[lane]):
            if peer:
                r.setSurfaces(peer_surfaces[k % 2], depths[lane])
                r.setColorLayout(n_samples == 1)   # tile-major slot (single sample only): two 128-byte lines per tile cross NVLink instead of eight 32-byte rows
            elif push:
                sink.before_render(k)
                r.setSurfaces(colors[k % 2], depths[lane])
            elif world > 1:
                gatherer.before_render(k)
                r.setSurfaces(colors[k % 2], depths[lane])
            else:
                r.setSurfaces(colors[k % len(colors)], depths[lane])
            r.setVertexBuffer(vb, 0)
            r.setIndexBuffer(ib, 0, n_tris)
            r.deferredClear((0.2, 0.4, 0.8, 1.0), 1.0)
            r.drawTriangles(asynchronous=asynchronous)
            if peer:
                sink.publish(k)
            elif push:
                sink.push(k)
            elif world > 1:
                gatherer.submit(k)
        return r

    def fork_lanes():
        for st in lane_streams[1:]:
            st.wait_stream(stream)

    def join_lanes():
        if gatherer or push:
            (gatherer or sink).finish()          # the last gather / copy is inside the timed region
        for st in lane_streams[1:]:
            stream.wait_stream(st)

    def finish_all():
        for r, st in zip(rasters, lane_streams):
            r.finish(stream=st.cuda_stream)

    def sync_all():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize(dev)

    # warm-up: synchronous draws size the work buffers (overflow-retry) and collect per-stage times
    stage_times = {s: [] for s in STAGES}
    for k in range(max(args.warmup, NUM_INPUT_COPIES) * F):
        st = step(k, asynchronous=False).getStats()
        for s, key in zip(STAGES, ("setupTime", "binTime", "coarseTime", "fineTime")):
            stage_times[s].append(st[key] * 1e3)
    for s in STAGES:   # drop the cold first frames
        stage_times[s] = stage_times[s][2:] or stage_times[s]
    join_lanes()
    sync_all()
    frame_ns_estimate = 1e6 * sum(statistics.median(v) for v in stage_times.values())
    launches_per_frame = raster.getLaunchCount()
    direct = raster.lastFrameDirect()   # automatic binning mode: small-triangle frames of an order-independent pipe skip the bin / coarse sort

    # The whole frame loop is ONE C call (crb_draw_batch_async) -- no Python between frames.  N > 1: every frame carries its
    # composite step (peer: the frame mark behind a render that went straight into rank 0's memory; push: DMA copy of the
    # finished frame into its slot on the library's side stream + mark); the NCCL composite keeps the Python loop.
    def make_batch(r, lane_depth):
        frames = []
        for k in range(args.steps):
            vb, ib = copies[k % NUM_INPUT_COPIES]
            fr = {"depth": lane_depth, "vb": vb, "ib": ib, "num_tris": n_tris, "clear": ((0.2, 0.4, 0.8, 1.0), 1.0)}
            if peer:
                fr.update(color=peer_surfaces[k % 2], signal_word=sink.mark_pointer(k), signal_value=k + 1)
            elif push:
                fr.update(color=colors[k % 2], push_dst=sink.slot_pointer(k), push_bytes=colors[k % 2].tensor.numel() * 4, slot=k % 2,
                          signal_word=sink.mark_pointer(k), signal_value=k + 1)
            else:
                fr.update(color=colors[k % len(colors)])
            frames.append(fr)
        return r.makeBatch(frames)
    batch_one = make_batch(raster, depth) if (world == 1 or peer or push) else None
    def timed_region(lanes, stage_events):
        """K frames enqueued back to back, bracketed by barrier + synchronize and two events on the main stream.  Returns
        (device ms per step, max over ranks; host ms spent enqueueing one step)."""
        raster.setStageTiming(stage_events)
        sync_all()
        e0.record(stream)
        if push and not args.no_stagger:
            # rank r starts r / N of a frame time late: the ranks' pushes then interleave at rank 0's NVLink ingress instead of arriving together
            crb.load_library().crb_ipc_delay(ctypes.c_void_p(stream.cuda_stream), int(rank * frame_ns_estimate / world))
        if lanes:
            fork_lanes()
        t0 = time.perf_counter()
        if batch_one is not None and not lanes:
            if peer:
                raster.setColorLayout(n_samples == 1)
            raster.drawBatch(batch_one)
            raster.batchJoin()                  # the side-stream copies are inside the timed region
        else:
            for k in range(args.steps):
                step(k, lanes=lanes)
        enqueue_ms = (time.perf_counter() - t0) * 1e3 / args.steps
        join_lanes()
        e1.record(stream)
        finish_all()
        sync_all()
        t = torch.tensor([e0.elapsed_time(e1)], dtype=torch.float64, device=dev)
        if world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item()) / args.steps, enqueue_ms

    # clocks are sampled on rank 0 only: NVML queries from every process of the job measurably slow down the CUDA calls of
    # all of them (enqueue_ms_per_step at N = 8: 0.14 ms with eight pollers, 0.03 ms without)
    sampler = ClockSampler(local) if rank == 0 else None
    if sampler:
        sampler.start()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    # ---- (1) THE METRIC: K frames, ONE frame in flight, the reference's five stage events on every frame
    bracket_ms, enqueue_ms = timed_region(lanes=False, stage_events=True)
    clocks = sampler.result() if sampler else None
    live = raster.getStageTiming()
    per_frame = raster.getStageTimingFrames()[-args.steps:]
    # t_frame of THIS rank: the four stage intervals of a frame; with a DMA composite on the side stream (push) the frame
    # pipeline has two stages -- render, copy of the previous frame -- and runs at the slower one
    stage_sum_ms = float(np.median(per_frame[:, :4].sum(axis=1))) if len(per_frame) else bracket_ms
    composite_ms = float(np.median(per_frame[:, 4])) if len(per_frame) else 0.0
    frame_ms = max(stage_sum_ms, composite_ms)
    raster.setStageTiming(False)
    per_rank = None
    if world > 1:
        t = torch.tensor([frame_ms], dtype=torch.float64, device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms_per_step = float(t.item())                            # the slowest rank's t_frame
        mine = torch.tensor([stage_sum_ms, composite_ms, enqueue_ms] + [float(np.median(per_frame[:, k])) for k in range(4)], dtype=torch.float64, device=dev)
        allr = [torch.zeros_like(mine) for _ in range(world)]
        dist.all_gather(allr, mine)
        per_rank = {"stage_sum_ms": [round(float(a[0]), 4) for a in allr], "composite_ms": [round(float(a[1]), 4) for a in allr], "enqueue_ms": [round(float(a[2]), 4) for a in allr],
                    "setup_ms": [round(float(a[3]), 4) for a in allr], "fine_ms": [round(float(a[6]), 4) for a in allr]}
    else:
        ms_per_step = frame_ms                                   # t_frame = sum of the four stage intervals, median over the K frames
    value = world * n_tris / (ms_per_step * 1e-3) / 1e6
    # ---- (2) the same K frames as one unbroken kernel chain (no stage events), and (3) alternating between two contexts
    chain_ms, chain_enqueue_ms = timed_region(lanes=False, stage_events=False)
    value_chain = world * n_tris / (chain_ms * 1e-3) / 1e6
    value_two = None
    if F > 1:
        two_ms, _ = timed_region(lanes=True, stage_events=False)
        value_two = world * n_tris / (two_ms * 1e-3) / 1e6

    composite_ok = None
    if peer or push:
        # outside the timed regions: every rank re-renders its last frame into a LOCAL surface; rank 0 compares the frames that
        # arrived in its memory over NVLink with them (checksums), and checks the per-slot frame marks
        k_last = args.steps - 1
        raster.setColorLayout(False)
        raster.setSurfaces(color, depth)
        vb, ib = copies[k_last % NUM_INPUT_COPIES]
        raster.setVertexBuffer(vb, 0)
        raster.setIndexBuffer(ib, 0, n_tris)
        raster.deferredClear((0.2, 0.4, 0.8, 1.0), 1.0)
        raster.drawTriangles()
        local_sum = color.tensor.to(torch.int64).sum().reshape(1)
        sums = [torch.zeros_like(local_sum) for _ in range(world)] if rank == 0 else None
        dist.gather(local_sum, sums, dst=0)
        sync_all()
        if rank == 0:
            marks = sink.read_marks(raster)
            composite_ok = bool((marks[k_last % 2] == k_last + 1).all())
            for r in range(world):
                got = sink.read_frame(raster, k_last, r, color.tensor.numel() * 4).view(np.int32).astype(np.int64).sum()
                composite_ok = composite_ok and int(got) == int(sums[r].item())
    if world > 1:
        raster.setColorLayout(False)
        raster.setSurfaces(color, depth)
    # ---- end to end through the host-buffer entry (pinned host memory in, colour frame out) ----------------
    h_color = torch.zeros_like(color.tensor, device="cpu").pin_memory()
    for _ in range(2):
        raster.deferredClear((0.2, 0.4, 0.8, 1.0), 1.0)
        raster.drawTrianglesHost(h_verts, h_idx, n_tris, h_color)
    sync_all()
    e2e_steps = max(3, min(args.steps, 10))
    e0.record(stream)
    for _ in range(e2e_steps):
        raster.deferredClear((0.2, 0.4, 0.8, 1.0), 1.0)
        raster.drawTrianglesHost(h_verts, h_idx, n_tris, h_color)
    e1.record(stream)
    sync_all()
    e2e_blocking_ms = e0.elapsed_time(e1) / e2e_steps
    # the same through the pipelined entry: every frame still uploads its inputs from pinned host memory and downloads its
    # colour frame, but upload(k+1) / render(k) / download(k-1) overlap; timed with events on the render stream, which the
    # last download is ordered before
    for _ in range(3):   # untimed: creates the copy streams and the two staging buffers
        raster.deferredClear((0.2, 0.4, 0.8, 1.0), 1.0)
        raster.drawTrianglesHostAsync(h_verts, h_idx, n_tris, h_color)
    raster.finish()
    sync_all()
    e0.record(stream)
    for _ in range(args.steps):
        raster.deferredClear((0.2, 0.4, 0.8, 1.0), 1.0)
        raster.drawTrianglesHostAsync(h_verts, h_idx, n_tris, h_color)
    e1.record(stream)
    raster.finish()
    sync_all()
    e2e_ms = torch.tensor([e0.elapsed_time(e1) / args.steps], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(e2e_ms, op=dist.ReduceOp.MAX)
    e2e_ms = float(e2e_ms.item())
    e2e_frame_ok = bool((h_color != 0).any().item())
    h2d = h_verts.numel() * 4 + h_idx.numel() * 4
    d2h = h_color.numel() * 4

    if rank == 0:
        med = {s: statistics.median(v) for s, v in stage_times.items()}   # synchronous warm-up frames (reference-style blocking draw)
        mean = {s: live[s] for s in STAGES}                               # asynchronous frames of the timed region, events on every frame
        line = {
            "metric": "Mtris/s", "value": value, "unit": "Mtris/s", "n_gpus": world, "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms_per_step,
            "frames_per_s": world * 1e3 / ms_per_step, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "s32/f32", "data": "synthetic",
            "config": config_dict(args.workload, world),
            "timing": "value = N * T / t_frame; t_frame = median over the K frames of the sum of the four stage intervals (CUDA events at the reference's five positions, SURVEY.md 8d), one frame in flight per rank"
                      + ("" if world == 1 else ", max over ranks; composite: " + ("inside the fine raster interval (the frame is rendered into rank 0's memory) + a frame mark" if peer else
                         "DMA copy of the finished frame on a side stream, overlapped with the next frame -- a rank's t_frame is the slower of its render (four intervals) and its copy (5th interval, composite_ms)" if push else
                         "NCCL gather on a side stream (not in t_frame: see bracket_ms_per_step)")),
            "composite_ms": composite_ms, "stage_sum_ms": stage_sum_ms, "per_rank": per_rank,
            "bracket_ms_per_step": bracket_ms, "enqueue_ms_per_step": enqueue_ms, "enqueue": ("one C call for the K frames, composite included (crb_draw_batch_async)" if batch_one is not None else "Python loop over crb_draw_triangles_async + NCCL gather calls"),
            "value_unbroken_chain": value_chain, "unbroken_chain_ms_per_step": chain_ms, "unbroken_chain_enqueue_ms_per_step": chain_enqueue_ms,
            "value_two_in_flight": value_two,
            "notes": {"binning": ("direct tile path: setup counts tiles -> queue allocation (timed as binRaster) -> unordered atomic scatter (timed as coarseRaster) -> fine raster keeps the (depth, index) minimum"
                                  if direct else "general path: stable two-level sort (bin raster, coarse raster), queues in submission order"),
                      "composite": None if world == 1 else (
                          "every rank renders straight into its frame slot in rank 0's memory (CUDA IPC peer memory over NVLink / NVSwitch): the composite is the render, no gather; frames verified on rank 0 after the timed region: %s" % composite_ok
                          if peer else "every rank renders locally and a DMA copy on a side stream (overlapped with the next frame) pushes the finished frame into its slot in rank 0's memory (CUDA IPC peer memory over NVLink / NVSwitch); frames verified on rank 0 after the timed region: %s" % composite_ok
                          if push else "every step's colour frames are gathered to rank 0 over NCCL inside the timed region (side stream, overlapped with the next frame's rendering)"),
                      "value_two_in_flight": "consecutive frames alternate between 2 contexts / streams / surface sets (double-buffered swap chain)",
                      "input_bytes_rotating": NUM_INPUT_COPIES * (verts.nbytes + idx.nbytes)},
            "stage_ms": mean, "device_frame_ms": sum(mean.values()), "device_frame_ms_median": frame_ms, "stage_ms_sync_draw": med, "stage_frames": live["frames"],
            "gpu_launches": (launches_per_frame + (1 if peer else 0)) * args.steps, "clocks": clocks,
            "e2e": {"value": world * n_tris / (e2e_ms * 1e-3) / 1e6, "unit": "Mtris/s", "ms_per_step": e2e_ms, "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h,
                    "api": "crb_draw_triangles_host_async (pinned host vertices+indices in, colour surface out; upload/render/download of consecutive frames overlap)",
                    "blocking_value": world * n_tris / (e2e_blocking_ms * 1e-3) / 1e6, "blocking_ms_per_step": e2e_blocking_ms,
                    "blocking_api": "crb_draw_triangles_host (one frame at a time, returns when the frame is on the host)", "frame_nonzero": e2e_frame_ok},
        }
        # ---- CPU oracle: baseline + algorithmic byte counts (bounded: one frame on all host threads) -------
        counts = None
        if not args.no_cpu_baseline and world == 1:
            from oracle import binding as G
            threads = G.hardware_threads()
            sample = n_tris    # one WHOLE frame: bounded (1-3 s on 8 threads) and it gives the counts of the roofline for every workload
            t, r = cpu_oracle_frame(verts, idx[:sample], w, h, shader, s_log2, flags, threads, True)
            line["cpu_baseline"] = {"value": sample / t / 1e6, "unit": "Mtris/s", "cores": threads, "kind": "port",
                                    "sample": "%d of %d triangles, one frame, oracle/golden.hpp on %d host threads (%.2f s)" % (sample, n_tris, threads, t)}
            counts = r["counts"]
        peak, peak_src = measured_peak()
        if counts is not None:
            sb = stage_bytes(counts, n_tris, k_var, ((w + 7) & ~7) * ((h + 7) & ~7), n_samples, bool(flags & 2))
            dom = max(STAGES, key=lambda s: mean[s])
            ach = sb[dom] / (mean[dom] * 1e-3) / 1e9
            line["roofline"] = {"bound": "hbm", "kernel": dom, "achieved": ach, "peak": peak, "unit": "GB/s", "frac": ach / peak, "traffic": None,
                                "peak_source": peak_src, "algorithmic_bytes_per_launch": sb[dom], "launch_ms": mean[dom]}
            b_alg = sum(sb.values())
            line["frame_roofline"] = {"algorithmic_bytes": b_alg, "achieved": b_alg / (sum(mean.values()) * 1e-3) / 1e9, "peak": peak, "unit": "GB/s",
                                      "frac": b_alg / (sum(mean.values()) * 1e-3) / 1e9 / peak,
                                      "stages": {s: {"bytes": sb[s], "ms": mean[s], "GB/s": sb[s] / (mean[s] * 1e-3) / 1e9} for s in STAGES}}
            prof = os.path.join(ROOT, "profiles", "traffic.json")
            if os.path.exists(prof):
                try:
                    line["roofline"]["traffic"] = json.load(open(prof)).get(args.workload, {}).get(dom)
                except Exception:
                    pass
        if not args.no_ref_kernels and world == 1:
            rk = time_ref_kernels(args.workload)
            if "Mtris/s" in rk:   # the north_star's comparison: same metric (T / sum of the four stage intervals, median), same frame, same GPU
                rk["speedup_of_this_pipeline"] = value / rk["Mtris/s"]
            line["ref_kernels"] = rk
        emit(line)
    if sink:
        sync_all()
        sink.close()
    for r in rasters:
        r.close()
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()
    return 0


if __name__ == "__main__":
    sys.exit(main())
