#!/usr/bin/env python3
"""Benchmark of the srp draw path on B200 (one JSON line on stdout, rank 0).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--workload cfg3]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N ... bench.py --gpus N ...

Workload (BASELINE.json: "frames/s & Gfrag/s (teapot 4K, 1M-tri mesh)"): cfg3, the synthetic
1 002 528-triangle shell at 3840x2160 with the camera inside the mesh, Gouraud shader, depth
test.  A step = srpFramebufferClear + srpDrawIndexBuffer of one frame per
GPU.  With N GPUs the headline is frame-parallel (every rank renders its own frames of the same
scene; rank 0 builds the mesh and broadcasts vertex / index buffers over NCCL once, no
collective in the timed region): weak scaling, value = N * K / max-over-ranks time.

  value      frames/s with inputs resident in HBM: device time of the K steps, measured with
             CUDA events on the library's own stream; L2 is flushed between steps
  e2e        the same metric through the public C API with HOST buffers: every step uploads
             the vertex and index buffers from pinned host memory (srp*BufferCopyData),
             clears, draws and brings the colour + depth planes back into the host-visible
             framebuffer; wall clock.  Headline: two frames in flight (explicit policy,
             srpB200FramebufferDownloadAsync / Wait); `synchronous`: the default policy, one
             frame at a time, every draw returning with the mirror up to date.  `host_link`
             puts it next to the measured ceiling of the host link (profiles/)
  roofline   the dominant kernel's algorithmic bytes / its measured duration vs the measured
             HBM copy bandwidth (MEASURED_PEAKS.json), plus the whole-frame figure
  parity     the benchmarked frame, at full size, against the unmodified reference (N = 1)
  secondary  cfg3 with heavy near-plane clipping (N = 1); cfg5 = BASELINE config 5, 1024 teapot
             frames of 1024x1024 split frame-parallel over the N ranks (srpB200DrawBatch)
  strips     (N > 1) ONE cfg3 frame split into N sort-first strips: every rank runs the geometry
             front-end and rasterises its rows straight into the root GPU's framebuffer over
             NVLink (CUDA IPC peer memory; flags in the root's memory signal completion), next
             to the NCCL send/recv gather as the baseline; bit-exactness vs one GPU is checked
  cpu_baseline  the unmodified reference (oracle/_ref) on the host cores, bounded sample

`--impl reference` times that reference instead (frame-parallel over all host cores).
The oracle is only used as the CPU baseline / reference arm / parity checker here, never on the
product path.
"""
from __future__ import annotations

import argparse
import ctypes as C
import json
import multiprocessing as mp
import os
import subprocess
import sys
import threading
import time
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parent
sys.path.insert(0, str(ROOT))

WORKLOADS = {
    "cfg3": dict(desc="cfg3: 1,002,528-triangle sphere shell, camera inside, 3840x2160, Gouraud VS, depth test",
                 width=3840, height=2160),
    "cfg2": dict(desc="cfg2: Utah teapot 1920x1080, Gouraud VS, depth test + back-face culling", width=1920, height=1080),
    "cfg1": dict(desc="cfg1: textured cube 800x600, depth test, perspective-correct UVs", width=800, height=600),
    # multi-draw / batch configs: timed on the GPU by tools/bench_all.py; here only the reference arm
    # (the CPU figures that stand next to them: profiles/r02_reference_arm_all_configs.jsonl)
    "cfg4": dict(desc="cfg4: 10 M sub-pixel triangles (two passes) + 1 M lines + 1 M points, stencil + scissor, 3840x2160", width=3840, height=2160,
                 reference_only=True),
    "cfg5": dict(desc="cfg5: one 1024x1024 frame of the teapot batch (frame-parallel: frames/s scale with the workers)", width=1024, height=1024,
                 reference_only=True),
}


def make_scene(workload):
    from srp_b200 import scenes as S
    if workload == "cfg3":
        return S.cfg3_shell()
    if workload == "cfg2":
        return S.cfg2_teapot()
    if workload == "cfg1":
        return S.cfg1_textured_cube()
    if workload == "cfg4":
        return S.cfg4_subpixel()
    if workload == "cfg5":
        return S.cfg5_frame(0)
    raise SystemExit(f"unknown workload {workload}")


def algorithmic_bytes(scene):
    """indices + referenced vertices + uniform + touched texels + one write of
    the three planes (9 B/px; the frame starts with a clear, so no read)"""
    d = scene.draws[0]
    n_idx = len(d.indices) if d.indices is not None else 0
    idx_bytes = n_idx * (d.indices.dtype.itemsize if d.indices is not None else 0)
    v_ref = len(np.unique(d.indices)) if d.indices is not None else d.vertices.nbytes // d.stride
    tex = sum(t[0].nbytes for t in scene.textures.values())
    uni = len(d.uniform) if isinstance(d.uniform, (bytes, bytearray)) else 200
    fb = scene.width * scene.height * 9
    return {"indices": idx_bytes, "vertices": v_ref * d.stride, "uniform": uni, "texture": tex, "framebuffer": fb,
            "total": idx_bytes + v_ref * d.stride + uni + tex + fb}


# ------------------------------------------------------------------------------------ clocks
class ClockSampler:
    QUERY = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
             "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, device):
        self.device, self.proc, self.lines = device, None, []

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--id={self.device}", f"--query-gpu={self.QUERY}",
                                          "--format=csv,noheader,nounits", "-lms", "20"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            threading.Thread(target=lambda: self.lines.extend(self.proc.stdout), daemon=True).start()
        except OSError:
            self.proc = None

    def stop(self):
        if self.proc is None:
            return None
        time.sleep(0.15)
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except subprocess.TimeoutExpired:
            self.proc.kill()
        sm, smax, reasons = [], [], set()
        for ln in self.lines:
            f = [x.strip() for x in ln.split(",")]
            if len(f) < 9:
                continue
            try:
                sm.append(float(f[1])); smax.append(float(f[2]))
            except ValueError:
                continue
            for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), f[5:9]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        if not sm:
            return None
        return {"sm_mhz": float(np.median(sm)), "sm_max_mhz": float(max(smax)), "reasons": sorted(reasons), "samples": len(sm)}


# ------------------------------------------------------------------------------------ CPU reference
def _ref_worker(args):
    workload, frames, warm = args
    from srp_b200 import scenes as S
    from oracle.refhost import load_oracle_reference      # the checker / baseline, never the product path
    ref = load_oracle_reference()
    p = S.Prepared(ref, make_scene(workload))
    for _ in range(warm):
        p.draw_all()
    t0 = time.perf_counter()
    for _ in range(frames):
        p.draw_all()
    dt = time.perf_counter() - t0
    covered = int((p.planes()[0] != 0).sum())
    p.free()
    return dt, covered


def run_reference(workload, frames_per_worker, workers, warm=1):
    """frame-parallel CPU run of the unmodified reference: `workers` processes (the library is
    single-threaded with global state), each renders `frames_per_worker` frames"""
    from oracle import refhost
    if not refhost.available():
        return None
    ctx = mp.get_context("spawn")
    t0 = time.perf_counter()
    with ctx.Pool(workers) as pool:
        res = pool.map(_ref_worker, [(workload, frames_per_worker, warm)] * workers)
    wall = time.perf_counter() - t0
    slowest = max(r[0] for r in res)
    return {"frames": frames_per_worker * workers, "seconds": slowest, "wall_with_setup": wall, "covered": res[0][1]}


def parity_vs_reference(lib, scene, label):
    """One frame of `scene` through the product (already the CUDA path) and through the unmodified
    reference, outside any timed region: are the three planes bit-identical?  The reference has no
    fragment counters, so the product's deterministic counts are reported, not compared."""
    from srp_b200 import scenes as S
    from oracle import refhost
    if not refhost.available():
        return {"config": label, "planes_equal": None, "why": "oracle/_ref not present on this box"}
    lib.dll.srpB200ResetStats()
    got = S.render(lib, scene)
    st = lib.stats()
    t0 = time.perf_counter()
    want = S.render(refhost.load_oracle_reference(), scene)
    ref_s = time.perf_counter() - t0
    diff = {n: int((a != b).sum()) for n, a, b in zip(("color", "depth", "stencil"), got, want)}
    return {"config": label, "planes_equal": all(v == 0 for v in diff.values()), "differing_pixels": diff,
            "pixels": int(got[0].size), "covered_pixels": int((want[0] != 0).sum()),
            "frags_emitted": st["fragsEmitted"], "frags_shaded": st["fragsShaded"], "prims_emitted": st["primsEmitted"],
            "frags_emitted_equal": None, "frags_note": "the unmodified reference exposes no fragment counter; planes are compared bit for bit",
            "reference_frame_s": ref_s, "checker": "unmodified reference (oracle/_ref), one frame, outside the timed region"}


def kernel_only_fps(lib, torch, stream, flush, scene, steps, warm):
    """device-timed frames/s of a secondary scene (same method as the headline value)"""
    from srp_b200 import host as H, scenes as S
    lib.dll.srpB200SetSyncMode(H.SRP_B200_SYNC_EXPLICIT)
    p = S.Prepared(lib, scene)
    for _ in range(warm):
        p.draw_all()
    lib.dll.srpB200Finish()
    lib.dll.srpB200ResetStats()
    ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(steps)]
    with torch.cuda.stream(stream):
        for a, b in ev:
            flush.fill_(3)
            a.record(stream); p.draw_all(); b.record(stream)
    torch.cuda.synchronize()
    ms = sum(a.elapsed_time(b) for a, b in ev) / steps
    st = lib.stats()
    p.free()
    lib.dll.srpB200SetSyncMode(H.SRP_B200_SYNC_DRAW)
    return {"frames_per_s": 1e3 / ms, "ms_per_frame": ms, "frags_emitted_per_frame": st["fragsEmitted"] / steps,
            "prims_emitted_per_frame": st["primsEmitted"] / steps, "prims_stored_per_frame": st["primsStored"] / steps}


def host_cores():
    try:
        return len(os.sched_getaffinity(0))
    except AttributeError:
        return os.cpu_count() or 1


# ------------------------------------------------------------------------------------ secondary: cfg5 batch
def cfg5_batch(lib, torch, dist, stream, world, rank, total_frames=1024, size=1024, steps=3):
    """BASELINE config 5: `total_frames` teapot frames of size x size, frame f rotated by f / 200,
    split frame-parallel over the ranks (multigpu.frame_partition), each rank's share in one
    srpB200DrawBatch call; device time of the call, max over ranks.  Returns rank-0's record."""
    from srp_b200 import host as H, multigpu as M, scenes as S
    mine = M.frame_partition(total_frames, world, rank)
    n = len(mine)
    mesh = S.teapot_mesh()
    draws = [S.teapot_draw(f, mesh) for f in mine]
    lib.new_context()
    for fn, *args in draws[0].state:
        getattr(lib.dll, fn)(*args)
    vb = lib.vertex_buffer(mesh[0], 32); ib = lib.index_buffer(mesh[1])
    prog = lib.program("gouraud", S.GOURAUD_VARYINGS, 12)
    fbs = [lib.framebuffer(size, size) for _ in range(n)]
    arr = (C.POINTER(H.SRPFramebuffer) * n)(*[f.ptr for f in fbs])
    uni = np.frombuffer(b"".join(d.uniform for d in draws), dtype=np.uint8).copy()
    stride = len(draws[0].uniform)
    prog.set_uniform(draws[0].uniform)
    lib.dll.srpB200SetSyncMode(H.SRP_B200_SYNC_EXPLICIT)
    call = lambda: lib.dll.srpB200DrawBatch(ib, vb, arr, n, C.byref(prog.sp), uni.ctypes.data, stride,
                                            H.SRP_PRIM_TRIANGLES, 0, len(mesh[1]), 1)
    call(); lib.dll.srpB200Finish()
    lib.dll.srpB200ResetStats(); lib.dll.srpB200SetProfiling(1); lib.stage_times()
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier(); torch.cuda.synchronize()
    ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(steps)]
    with torch.cuda.stream(stream):
        for a, b in ev:
            a.record(stream); call(); b.record(stream)
    torch.cuda.synchronize()
    ms = sum(a.elapsed_time(b) for a, b in ev) / steps
    st = lib.stage_times(); lib.dll.srpB200SetProfiling(0)
    stats = lib.stats()
    # one frame of the batch against a single clear + draw of the same frame (this process, same library)
    probe = fbs[n // 2].planes()
    lib.dll.srpB200SetSyncMode(H.SRP_B200_SYNC_DRAW)
    single = S.render(lib, S.cfg5_frame(mine[n // 2], size, mesh))
    same = all(bool(np.array_equal(a, b)) for a, b in zip(probe, single))
    for f in fbs:
        f.free()
    lib.dll.srpFreeVertexBuffer(vb); lib.dll.srpFreeIndexBuffer(ib)
    if world > 1:
        t = torch.tensor([ms, 0.0 if same else 1.0], dtype=torch.float64, device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms, same = float(t[0]), float(t[1]) == 0.0
    per_frame_bytes = size * size * 9 + mesh[0].nbytes + mesh[1].nbytes
    return {"what": f"{total_frames} teapot frames {size}x{size} ({mesh[2]}), frame-parallel over {world} rank(s), one srpB200DrawBatch per rank",
            "frames_per_s": total_frames / ms * 1e3, "ms_per_batch": ms, "frames_per_rank": n,
            "stage_ms_per_batch_rank0": {k: st[k] / steps for k in ("geometry_ms", "binning_ms", "tiles_ms")},
            "frags_emitted_per_frame_rank0": stats["fragsEmitted"] / steps / n,
            "hbm_gbs_per_gpu_algorithmic": n * per_frame_bytes / ms / 1e6,
            "batch_frame_equals_single_draw": same}


# ------------------------------------------------------------------------------------ strips (N > 1)
def strips(lib, torch, dist, stream, world, rank, prep, scene, steps, warm, single_planes):
    """ONE frame split into `world` sort-first strips.  (a) fused: the other ranks' tile kernels
    write their strips into the root's framebuffer over NVLink (multigpu.StripTarget);
    (b) baseline: every rank renders locally, NCCL send/recv gathers the strips.  Device time on
    the root from the first frame's submission to the last frame being complete in its memory."""
    from srp_b200 import host as H, multigpu as M
    lib.dll.srpB200SetSyncMode(H.SRP_B200_SYNC_EXPLICIT)
    keep = prep.fb
    out = {"ranks": world, "rows_per_rank": [list(M.strip_rows(scene.height, int(lib.dll.srpB200TileHeight()), world, r)) for r in range(world)]}

    def draw_into(fb):
        prep.fb = fb
        prep.draw_all()

    # (a) fused peer write, consecutive frames on alternating lanes (the replicated geometry
    # front-end of frame k+1 runs under frame k's tiles)
    strip_lanes = max(1, min(4, int(lib.dll.srpB200LaneCount())))
    target = M.StripTarget(lib, scene.width, scene.height, ring=strip_lanes, root=0, lanes=strip_lanes)
    lane_streams = []
    for l in range(strip_lanes):
        lib.dll.srpB200SetLane(l)
        lane_streams.append(torch.cuda.ExternalStream(lib.dll.srpB200Stream()))
    lib.dll.srpB200SetLane(0)
    def run(n):
        for _ in range(n):
            target.render(draw_into)
            if rank == 0:
                target.complete()
    run(max(warm, 2 * strip_lanes))
    lib.dll.srpB200Finish()
    torch.cuda.synchronize(); dist.barrier(); torch.cuda.synchronize()
    s_ev = [torch.cuda.Event(enable_timing=True) for _ in lane_streams]
    e_ev = [torch.cuda.Event(enable_timing=True) for _ in lane_streams]
    for ev, st_ in zip(s_ev, lane_streams):
        ev.record(st_)
    run(steps)
    for ev, st_ in zip(e_ev, lane_streams):
        ev.record(st_)
    lib.dll.srpB200Finish()
    torch.cuda.synchronize()
    ms = max(a.elapsed_time(b) for a in s_ev for b in e_ev)
    t = torch.tensor([ms], dtype=torch.float64, device="cuda")
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    exact = None
    if rank == 0:
        last = target.fbs[(target.frame - 1) % target.ring]
        got = last.planes()
        exact = all(bool(np.array_equal(a, b)) for a, b in zip(got, single_planes))
    dist.barrier()
    peer_bytes = 0 if rank == 0 else (target.rows[1] - target.rows[0]) * scene.width * 8      # colour + depth rows written remotely
    tb = torch.tensor([float(peer_bytes)], dtype=torch.float64, device="cuda")
    dist.all_reduce(tb)
    target.free()
    out["fused_peer_write"] = {"frames_per_s": steps / (float(t[0]) / 1e3), "ms_per_frame": float(t[0]) / steps,
                               "bit_exact_vs_single_gpu": exact, "nvlink_bytes_per_frame": int(tb[0]),
                               "what": "tile kernels of ranks 1.. store their strips into rank 0's planes (CUDA IPC peer memory) as part of the tile "
                                       f"write-back; per frame and rank one flag store in rank 0's memory signals completion; {strip_lanes} frames in flight (lanes), one framebuffer each", "frames_in_flight": strip_lanes}

    # (b) NCCL gather baseline
    th = int(lib.dll.srpB200TileHeight())
    r0, r1 = M.strip_rows(scene.height, th, world, rank)
    fb = lib.framebuffer(scene.width, scene.height)
    planes = [M.device_plane_tensor(lib, fb, w) for w in range(3)]
    def run_gather(n):
        for _ in range(n):
            lib.dll.srpB200SetRowRange(r0, r1)
            draw_into(fb)
            lib.dll.srpB200SetRowRange(0, 2 ** 64 - 1)
            with torch.cuda.stream(stream):
                M.gather_strips_inplace(planes, scene.height, th, dst=0)
    run_gather(warm)
    lib.dll.srpB200Finish(); torch.cuda.synchronize(); dist.barrier(); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    with torch.cuda.stream(stream):
        e0.record(stream)
    run_gather(steps)
    with torch.cuda.stream(stream):
        e1.record(stream)
    lib.dll.srpB200Finish(); torch.cuda.synchronize()
    t = torch.tensor([e0.elapsed_time(e1)], dtype=torch.float64, device="cuda")
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    exact2 = None
    if rank == 0:
        exact2 = all(bool(np.array_equal(a, b)) for a, b in zip(fb.planes(), single_planes))
    fb.free()
    out["nccl_gather"] = {"frames_per_s": steps / (float(t[0]) / 1e3), "ms_per_frame": float(t[0]) / steps,
                          "bit_exact_vs_single_gpu": exact2,
                          "what": "every rank renders its strip locally; one batch of ncclSend / ncclRecv per frame moves the strips of the three planes to rank 0"}
    prep.fb = keep
    lib.dll.srpB200SetSyncMode(H.SRP_B200_SYNC_DRAW)
    return out


# ------------------------------------------------------------------------------------ main
def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default="cfg3", choices=sorted(WORKLOADS))
    ap.add_argument("--lanes", type=int, default=4, help="frames in flight in the headline measurement (srpB200SetLane); 1 = one frame at a time")
    ap.add_argument("--cpu-seconds", type=float, default=12.0, help="budget of the cpu_baseline sample (0: skip it and the secondary legs, e.g. under ncu)")
    args = ap.parse_args()

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    W = max(args.warmup, 3)
    K = args.steps
    wl = WORKLOADS[args.workload]
    # the same dict in both arms (the driver compares them): what differs between the arms -- GPUs
    # vs host processes -- is stated under n_gpus / cpu_baseline.cores
    lanes = max(1, min(args.lanes, 4))
    config = {"workload": wl["desc"], "frames_per_step_per_worker": 1,
              "partition": "frame-parallel: every worker (GPU rank / host process) renders its own frames of the same scene",
              "frames_in_flight": f"ours: {lanes} per GPU (consecutive frames on alternating lanes = streams with their own scratch pools, srpB200SetLane); "
                                  "reference arm: one per host process",
              "l2": (f"ours: inputs larger than L2 -- the frames rotate over {2 * lanes} independent sets of vertex buffer, index buffer and framebuffer "
                     "(no flush inside the timed region); the one-frame-at-a-time figures (stage_ms_per_frame, roofline, one_frame_in_flight) are timed per "
                     "step with a 256 MiB fill between steps" if lanes > 1 else "ours: flushed between timed steps (256 MiB fill)") + "; reference arm: n/a (host caches)",
              "data": "synthetic"}

    if args.impl == "reference":
        if rank != 0:
            return
        from oracle import refhost
        if not refhost.available():
            print(json.dumps({"impl": "reference", "unavailable": "oracle/_ref/libref_host.so was not built (needs /root/reference at build time)"}))
            return
        cores = host_cores()
        workers = min(cores, 64)
        r = run_reference(args.workload, K, workers, warm=min(W, 1))
        value = r["frames"] / r["seconds"]
        line = {"impl": "reference", "metric": f"frames_per_s_{args.workload}", "value": value, "unit": "frames/s",
                "n_gpus": args.gpus, "steps": K, "warmup": W, "ms_per_step": 1e3 * r["seconds"] / K,
                "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
                "config": config,
                "cpu_baseline": {"value": value, "unit": "frames/s", "cores": workers, "kind": "reference",
                                 "sample": f"{r['frames']} frames of the same workload ({K} per process), unmodified reference built by oracle/Makefile"},
                "e2e": {"value": value, "unit": "frames/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
                "gpu_launches": 0}
        print(json.dumps(line))
        return

    # ---------------------------------------------------------------- ours
    if wl.get("reference_only"):
        raise SystemExit(f"--workload {args.workload} is a multi-draw / batch config: its GPU timings come from tools/bench_all.py; "
                         "bench.py runs it with --impl reference only")
    os.environ["SRP_B200_DEVICE"] = str(local_rank)
    # NUMA placement of this rank's host side BEFORE anything pins memory: the pinned framebuffer
    # mirrors and staging buffers then live next to the rank's GPU (srp_b200/numa.py)
    from srp_b200 import numa
    numa_note = numa.bind_to_gpu(local_rank) if os.environ.get("SRP_B200_NUMA_BIND", "1") != "0" else "off (SRP_B200_NUMA_BIND=0)"
    import torch
    import torch.distributed as dist
    torch.cuda.set_device(local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    from srp_b200 import host as H, scenes as S
    lib = H.load_product()
    scene = make_scene(args.workload)
    draw = scene.draws[0]
    alg = algorithmic_bytes(scene)

    # rank 0's vertex / index buffers are broadcast over NVLink once; every rank then uploads
    # from the device copy (srp*BufferCopyData accepts device memory through unified addressing)
    vraw = np.ascontiguousarray(draw.vertices).view(np.uint8).reshape(-1)
    iraw = np.ascontiguousarray(draw.indices).view(np.uint8).reshape(-1)
    vpin = torch.from_numpy(vraw.copy()).pin_memory()
    ipin = torch.from_numpy(iraw.copy()).pin_memory()
    if world > 1:
        vdev, idev = vpin.cuda(), ipin.cuda()
        if rank != 0:
            vdev.zero_(); idev.zero_()
        dist.broadcast(vdev, 0); dist.broadcast(idev, 0)
        torch.cuda.synchronize()

    prep = S.Prepared(lib, scene)
    _, prog, vb, ib, count = prep.items[0]
    if world > 1:
        lib.dll.srpVertexBufferCopyData(vb, draw.stride, vdev.numel(), vdev.data_ptr())
        lib.dll.srpIndexBufferCopyData(ib, H.SRP_UINT32, idev.numel(), idev.data_ptr())
    stream = torch.cuda.ExternalStream(lib.dll.srpB200Stream())
    flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")

    def barrier():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
            torch.cuda.synchronize()

    # ---- kernel-only: inputs resident, explicit synchronisation, CUDA events on the library stream
    lib.dll.srpB200SetSyncMode(H.SRP_B200_SYNC_EXPLICIT)
    for _ in range(W):
        prep.draw_all()
    lib.dll.srpB200Finish()
    lib.dll.srpB200ResetStats()
    lib.dll.srpB200SetProfiling(1)
    lib.stage_times()
    sampler = ClockSampler(local_rank)
    barrier()
    sampler.start()
    starts = [torch.cuda.Event(enable_timing=True) for _ in range(K)]
    ends = [torch.cuda.Event(enable_timing=True) for _ in range(K)]
    with torch.cuda.stream(stream):
        for k in range(K):
            flush.fill_(k & 0xFF)                 # L2 flush, outside the timed interval
            starts[k].record(stream)
            prep.draw_all()
            ends[k].record(stream)
    barrier()
    dev_ms = sum(s.elapsed_time(e) for s, e in zip(starts, ends))
    stages = lib.stage_times()
    lib.dll.srpB200SetProfiling(0)
    stats = lib.stats()
    frags_per_frame = stats["fragsEmitted"] / max(1, stats["draws"])
    shaded_per_frame = stats["fragsShaded"] / max(1, stats["draws"])
    launches = stats["kernelLaunches"]

    # ---- frames in flight (the headline when lanes > 1): the same K frames, consecutive frames on
    # alternating lanes, so the geometry / binning kernels of one frame (which leave most of the
    # SMs idle) run under the tile kernel of another.  Timed as ONE interval over all K steps with
    # events on every lane's stream (latest end - earliest start); nothing is flushed inside the
    # interval, instead the frames rotate over 2 * lanes independent sets of buffers + framebuffer
    # (cfg3: 103 MB each), more than the 126 MB L2 holds.
    flight_ms = None
    ring = [prep]
    if lanes > 1:
        lanes = min(lanes, int(lib.dll.srpB200LaneCount()))
        ring += [S.Prepared(lib, scene) for _ in range(2 * lanes - 1)]
        lane_streams = []
        for l in range(lanes):
            lib.dll.srpB200SetLane(l)
            lane_streams.append(torch.cuda.ExternalStream(lib.dll.srpB200Stream()))
        lib.dll.srpB200SetLane(0)
        def flight_run(n):
            for k in range(n):
                lib.dll.srpB200SetLane(k % lanes)
                ring[k % len(ring)].draw_all()
            lib.dll.srpB200SetLane(0)
        flight_run(max(W, 2 * len(ring)))
        lib.dll.srpB200Finish()
        lib.dll.srpB200ResetStats()
        barrier()
        s_ev = [torch.cuda.Event(enable_timing=True) for _ in range(lanes)]
        e_ev = [torch.cuda.Event(enable_timing=True) for _ in range(lanes)]
        for l in range(lanes):
            s_ev[l].record(lane_streams[l])
        flight_run(K)
        for l in range(lanes):
            e_ev[l].record(lane_streams[l])
        barrier()
        flight_ms = max(a.elapsed_time(b) for a in s_ev for b in e_ev)
        launches = lib.stats()["kernelLaunches"]
        first = ring[0].planes()
        flight_equal = all(bool(np.array_equal(x, y)) for p in ring[1:] for x, y in zip(p.planes(), first))
        del first
        for p in ring[1:]:
            p.free()
    clocks = sampler.stop()

    # ---- end to end: host buffers in, host-visible framebuffer out, default policy, wall clock
    lib.dll.srpB200SetSyncMode(H.SRP_B200_SYNC_DRAW)
    def e2e_step():
        lib.dll.srpVertexBufferCopyData(vb, draw.stride, vpin.numel(), vpin.data_ptr())
        lib.dll.srpIndexBufferCopyData(ib, H.SRP_UINT32, ipin.numel(), ipin.data_ptr())
        prep.draw_all()                           # returns with colour + depth mirrored on the host
    for _ in range(W):
        e2e_step()
    lib.dll.srpB200ResetStats()
    barrier()
    t0 = time.perf_counter()
    for _ in range(K):
        e2e_step()
    torch.cuda.synchronize()
    e2e_s = time.perf_counter() - t0
    barrier()
    st2 = lib.stats()
    # two friendlier variants of the same step, reported next to the headline e2e figure:
    # (a) the mesh stays resident (real srp programs upload once and change only the uniform),
    # (b) additionally only the colour plane is mirrored (what the reference's own programs read)
    def timed(fn):
        for _ in range(W):
            fn()
        barrier()
        t = time.perf_counter()
        for _ in range(K):
            fn()
        torch.cuda.synchronize()
        dt = time.perf_counter() - t
        barrier()
        return dt
    e2e_static_s = timed(prep.draw_all)
    # (c) pipelined: the same per-step work (mesh upload from pinned memory, clear, draw, colour +
    # depth back in host memory) under the explicit policy with TWO framebuffers in flight:
    # step i's planes cross PCIe (srpB200FramebufferDownloadAsync) while step i+1 uploads and
    # renders; the host waits for step i-1's mirror and reads it before it enqueues step i+1's
    # successor.  Every step's H2D and D2H copies lie inside the timed region.
    lib.dll.srpB200SetSyncMode(H.SRP_B200_SYNC_EXPLICIT)
    fb2 = lib.framebuffer(scene.width, scene.height)
    fbs = [prep.fb, fb2]
    pipe_checks = []
    def pipe_run(n):
        for i in range(n):
            fb = fbs[i & 1]
            lib.dll.srpVertexBufferCopyData(vb, draw.stride, vpin.numel(), vpin.data_ptr())
            lib.dll.srpIndexBufferCopyData(ib, H.SRP_UINT32, ipin.numel(), ipin.data_ptr())
            prep.fb = fb
            prep.draw_all()
            lib.dll.srpB200FramebufferDownloadAsync(fb.ptr)
            if i > 0:
                prev = fbs[(i - 1) & 1]
                lib.dll.srpB200FramebufferWait(prev.ptr)
                pipe_checks.append(int(prev.ptr.contents.color[(scene.height // 2) * scene.width + scene.width // 2]))
        last = fbs[(n - 1) & 1]
        lib.dll.srpB200FramebufferWait(last.ptr)
        pipe_checks.append(int(last.ptr.contents.color[(scene.height // 2) * scene.width + scene.width // 2]))
    pipe_run(W)
    lib.dll.srpB200Finish()
    lib.dll.srpB200ResetStats()
    barrier()
    t0 = time.perf_counter()
    pipe_run(K)
    lib.dll.srpB200Finish()
    e2e_pipe_s = time.perf_counter() - t0
    barrier()
    st3 = lib.stats()
    prep.fb = fbs[0]
    pipe_checksum = int(np.ctypeslib.as_array(fbs[(K - 1) & 1].ptr.contents.color, shape=(scene.width * scene.height,))[::4099].sum())
    fb2.free()
    lib.dll.srpB200SetSyncMode(H.SRP_B200_SYNC_DRAW)
    lib.dll.srpB200SetMirrorPlanes(1)
    e2e_color_s = timed(prep.draw_all)
    lib.dll.srpB200SetMirrorPlanes(7)
    single_planes = prep.fb.planes()
    checksum = int(single_planes[0].reshape(-1)[::4099].sum())

    if world > 1:
        t = torch.tensor([dev_ms, e2e_s * 1e3, e2e_static_s * 1e3, e2e_color_s * 1e3, e2e_pipe_s * 1e3, flight_ms or 0.0], dtype=torch.float64, device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        dev_ms, e2e_s, e2e_static_s, e2e_color_s, e2e_pipe_s = float(t[0]), float(t[1]) / 1e3, float(t[2]) / 1e3, float(t[3]) / 1e3, float(t[4]) / 1e3
        if flight_ms is not None:
            flight_ms = float(t[5])

    # ---- secondary legs that need every rank
    secondary = {}
    strips_rec = None
    full = args.cpu_seconds > 0
    if full and world > 1 and args.workload == "cfg3":
        try:
            strips_rec = strips(lib, torch, dist, stream, world, rank, prep, scene, K, W, single_planes)
        except Exception as e:      # noqa: BLE001 -- a secondary leg must not cost the headline line
            strips_rec = {"error": f"{type(e).__name__}: {e}"}
    if full:
        try:
            secondary["cfg5_batch_frame_parallel"] = cfg5_batch(lib, torch, dist, stream, world, rank)
        except Exception as e:      # noqa: BLE001
            secondary["cfg5_batch_frame_parallel"] = {"error": f"{type(e).__name__}: {e}"}

    if rank == 0:
        peaks = {}
        pk = ROOT / "MEASURED_PEAKS.json"
        if pk.exists():
            peaks = json.loads(pk.read_text())
        hbm = float(peaks.get("hbm_gbs", 6650.0))
        peak_src = "measured (MEASURED_PEAKS.json hbm_gbs)" if "hbm_gbs" in peaks else "fallback 6.65 TB/s (B200_PROFILING.md)"
        one_at_a_time = world * K / (dev_ms / 1e3)
        head_ms = flight_ms if flight_ms is not None else dev_ms      # the K timed steps of the headline
        value = world * K / (head_ms / 1e3)
        n = max(1, stages["draws"])
        per = {k: stages[k] / n for k in ("geometry_ms", "binning_ms", "tiles_ms")}
        # dominant kernel and its algorithmic bytes per launch (DESIGN.md "Kernels"):
        #   geometry: indices + referenced vertices in; tiles: one write of the three planes
        dom = max(per, key=per.get)
        dom_bytes = {"geometry_ms": alg["indices"] + alg["vertices"] + alg["uniform"], "binning_ms": 0,
                     "tiles_ms": alg["framebuffer"] + alg["texture"]}[dom]
        achieved = dom_bytes / (per[dom] / 1e3) / 1e9 if per[dom] > 0 else 0.0
        frame_gbs = alg["total"] * (value / world) / 1e9
        traffic = None
        tr = ROOT / "profiles" / "r02_cfg3_ncu_summary_traffic.json"
        dom_kernel = {"geometry_ms": "srpdGeomKernel", "binning_ms": "srpdBinFillKernel", "tiles_ms": "srpdTileKernel"}[dom]
        if tr.exists() and args.workload == "cfg3":
            traffic = json.loads(tr.read_text())["kernels"].get(dom_kernel, {}).get("dram_bytes")
        ceiling = None
        cf = ROOT / "profiles" / f"r02_pcie_ceiling_n{world}.json"
        if cf.exists():
            c = json.loads(cf.read_text())
            ceiling = {"frames_per_s_ceiling": c["frames_per_s_ceiling_aggregate"], "d2h_gbs_per_rank": c["d2h_alone_gbs_per_rank"],
                       "h2d_gbs_per_rank": c["h2d_alone_gbs_per_rank"], "both_gbs_per_rank": c["both_gbs_per_rank"],
                       "fraction_of_ceiling": (world * K / e2e_pipe_s) / c["frames_per_s_ceiling_aggregate"],
                       "source": f"profiles/r02_pcie_ceiling_n{world}.json: pinned copies of one frame's traffic (28 MB up, 66 MB down) on every rank at once"}
        line = {
            "metric": f"frames_per_s_{args.workload}", "value": value, "unit": "frames/s",
            "n_gpus": world, "steps": K, "warmup": W, "ms_per_step": head_ms / K,
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": config,
            "frames_in_flight": lanes,
            "one_frame_in_flight": {"value": one_at_a_time, "ms_per_step": dev_ms / K,
                                    "what": "the same K frames one at a time on one lane, each step timed by its own pair of events, L2 flushed between steps; "
                                            "stage_ms_per_frame and roofline come from this run"},
            "gfrag_per_s": world * frags_per_frame * K / (head_ms / 1e3) / 1e9,
            "fragments_per_frame": frags_per_frame, "shaded_fragments_per_frame": shaded_per_frame,
            "input_triangles_per_s": world * (count // 3) * K / (head_ms / 1e3),
            "stage_ms_per_frame": per,
            "roofline": {"bound": "hbm", "kernel": {"geometry_ms": "srpdGeomKernel", "binning_ms": "srpdBin*Kernel", "tiles_ms": "srpdTileKernel"}[dom],
                         "achieved": achieved, "peak": hbm, "unit": "GB/s", "frac": achieved / hbm, "traffic": traffic,
                         "traffic_source": "profiles/r02_cfg3_ncu_summary_traffic.json (ncu --set full, one full-frame launch; the 126 MB L2 absorbs part of the plane writes within the launch)" if traffic else None,
                         "algorithmic_bytes_per_launch": dom_bytes, "peak_source": peak_src,
                         "frame": {"algorithmic_bytes": alg, "achieved": frame_gbs, "frac": frame_gbs / hbm}},
            "e2e": {"value": world * K / e2e_pipe_s, "unit": "frames/s",
                    "h2d_bytes_per_step": st3["h2dBytes"] // K, "d2h_bytes_per_step": st3["d2hBytes"] // K,
                    "ms_per_step": 1e3 * e2e_pipe_s / K, "result_checksum": pipe_checksum,
                    "result_checksum_matches_synchronous": pipe_checksum == checksum and len(set(pipe_checks)) == 1,
                    "host_link": ceiling, "numa": numa_note,
                    "what": "throughput with two frames in flight through the public C API (explicit synchronisation policy): per step "
                            "srpVertexBufferCopyData + srpIndexBufferCopyData from pinned host memory, srpFramebufferClear, "
                            "srpDrawIndexBuffer, srpB200FramebufferDownloadAsync (colour + depth into the host-visible framebuffer), and "
                            "srpB200FramebufferWait + a host read of the previous step's planes; every step's copies are inside the timed region",
                    "synchronous": {"value": world * K / e2e_s, "ms_per_step": 1e3 * e2e_s / K,
                                    "h2d_bytes_per_step": st2["h2dBytes"] // K, "d2h_bytes_per_step": st2["d2hBytes"] // K,
                                    "what": "default policy, one frame at a time: every srpDrawIndexBuffer returns with colour + depth mirrored on the host"},
                    "variants": {"mesh_resident_frames_per_s": world * K / e2e_static_s,
                                 "mesh_resident_color_only_frames_per_s": world * K / e2e_color_s}},
            "gpu_launches": launches,
            "clocks": clocks,
            "version": lib.dll.srpB200Version().decode(),
        }
        if flight_ms is not None:
            line["frames_in_flight_planes_equal"] = flight_equal      # every set's framebuffer == the one-at-a-time render
        if strips_rec is not None:
            line["strips"] = strips_rec
        if world == 1 and full:
            # parity of the benchmarked workload itself, at full size, against the unmodified reference
            line["parity"] = parity_vs_reference(lib, scene, wl["desc"])
            if args.workload == "cfg3":
                # BASELINE config 3 says "heavy near-plane clipping": the r = 3 shell of the headline
                # clips 0.2 % of its triangles, so the r = 1.2 shell (every triangle near the camera
                # straddles the near plane) is timed and checked as well
                heavy = S.cfg3_shell(radius=1.2)
                sec = kernel_only_fps(lib, torch, stream, flush, heavy, K, W)
                sec["parity"] = parity_vs_reference(lib, heavy, "cfg3 with shell radius 1.2 (heavy near-plane clipping)")
                secondary["cfg3_r1.2_heavy_clipping"] = sec
            cores = host_cores()
            workers = min(cores, 64)
            per_frame_guess = 0.6 if args.workload == "cfg3" else 0.03
            fpw = max(1, int(args.cpu_seconds / per_frame_guess))
            r = run_reference(args.workload, fpw, workers)
            if r is not None:
                line["cpu_baseline"] = {"value": r["frames"] / r["seconds"], "unit": "frames/s", "cores": workers, "kind": "reference",
                                        "single_core_value": (r["frames"] / workers) / r["seconds"],
                                        "sample": f"{r['frames']} frames of the same workload ({fpw} per process, {workers} single-threaded processes), "
                                                  "unmodified reference built by oracle/Makefile"}
            else:
                line["cpu_baseline"] = {"value": None, "unit": "frames/s", "cores": 0, "kind": "reference",
                                        "sample": "oracle/_ref not present on this box"}
        if secondary:
            line["secondary"] = secondary
        print(json.dumps(line))
    prep.free()
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
