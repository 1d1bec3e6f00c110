#!/usr/bin/env python3
"""Secondary measurements (not the driver's JSON line): kernel-only frame times and stage
split of every BASELINE config on one GPU, incl. the frame-parallel batch (cfg5).
Run under gpurun; writes gpurun_out/bench_all.json.  usage: bench_all.py [cfg1 cfg2 cfg2_4k cfg3 cfg3_r12 cfg4 cfg5_frame cfg5 ...]"""
import ctypes as C
import json
import sys
import time
from pathlib import Path

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
import numpy as np
import torch
from srp_b200 import host as H, scenes as S


def time_scene(lib, scene, steps=10, warm=3):
    lib.dll.srpB200SetSyncMode(H.SRP_B200_SYNC_EXPLICIT)
    p = S.Prepared(lib, scene)
    for _ in range(warm):
        p.draw_all()
    lib.dll.srpB200Finish()
    lib.dll.srpB200ResetStats(); lib.dll.srpB200SetProfiling(1); lib.stage_times()
    stream = torch.cuda.ExternalStream(lib.dll.srpB200Stream())
    flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
    ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(steps)]
    with torch.cuda.stream(stream):
        for a, b in ev:
            flush.fill_(1)
            a.record(stream); p.draw_all(); b.record(stream)
    torch.cuda.synchronize()
    ms = sum(a.elapsed_time(b) for a, b in ev) / steps
    st = lib.stage_times(); lib.dll.srpB200SetProfiling(0)
    stats = lib.stats()
    t0 = time.perf_counter()
    lib.dll.srpB200SetSyncMode(H.SRP_B200_SYNC_DRAW)
    for _ in range(steps):
        p.draw_all()
    e2e = (time.perf_counter() - t0) / steps * 1e3
    p.free()
    n = max(1, st["draws"])
    return {"scene": scene.name, "ms_per_frame": ms, "frames_per_s": 1e3 / ms, "e2e_ms_per_frame_sync_draw": e2e,
            "draws_per_frame": len(scene.draws), "stage_ms_per_draw": {k: st[k] / n for k in ("geometry_ms", "binning_ms", "tiles_ms")},
            "frags_emitted_per_frame": stats["fragsEmitted"] / steps, "frags_shaded_per_frame": stats["fragsShaded"] / steps,
            "gfrag_per_s": stats["fragsEmitted"] / steps / ms / 1e6, "launches_per_frame": stats["kernelLaunches"] / steps}


def time_in_flight(lib, scene, lanes=4, steps=200):
    """the same frames with `lanes` of them in flight (srpB200SetLane): one interval over all steps,
    events on every lane's stream, one Prepared (buffers + framebuffer) per lane"""
    lib.dll.srpB200SetSyncMode(H.SRP_B200_SYNC_EXPLICIT)
    lanes = min(lanes, int(lib.dll.srpB200LaneCount()))
    ring = [S.Prepared(lib, scene) for _ in range(lanes)]
    streams = []
    for l in range(lanes):
        lib.dll.srpB200SetLane(l)
        streams.append(torch.cuda.ExternalStream(lib.dll.srpB200Stream()))
    def run(n):
        for k in range(n):
            lib.dll.srpB200SetLane(k % lanes)
            ring[k % lanes].draw_all()      # (every frame sets the same state: no fresh context needed in between)
        lib.dll.srpB200SetLane(0)
    run(2 * lanes)
    lib.dll.srpB200Finish()
    s_ev = [torch.cuda.Event(enable_timing=True) for _ in streams]
    e_ev = [torch.cuda.Event(enable_timing=True) for _ in streams]
    for ev, st in zip(s_ev, streams):
        ev.record(st)
    run(steps)
    for ev, st in zip(e_ev, streams):
        ev.record(st)
    lib.dll.srpB200Finish()
    torch.cuda.synchronize()
    ms = max(a.elapsed_time(b) for a in s_ev for b in e_ev) / steps
    first = ring[0].planes()
    same = all(bool(np.array_equal(x, y)) for p in ring[1:] for x, y in zip(p.planes(), first))
    for p in ring:
        p.free()
    lib.dll.srpB200SetSyncMode(H.SRP_B200_SYNC_DRAW)
    return {"frames_in_flight": lanes, "ms_per_frame": ms, "frames_per_s": 1e3 / ms, "planes_equal_across_lanes": same}


def time_batch(lib, n_frames=1024, size=1024, steps=3):
    mesh = S.teapot_mesh()
    draws = [S.teapot_draw(f, mesh) for f in range(n_frames)]
    lib.new_context()
    for fn, *args in draws[0].state:
        getattr(lib.dll, fn)(*args)
    vb = lib.vertex_buffer(mesh[0], 32); ib = lib.index_buffer(mesh[1])
    prog = lib.program("gouraud", S.GOURAUD_VARYINGS, 12)
    fbs = [lib.framebuffer(size, size) for _ in range(n_frames)]
    arr = (C.POINTER(H.SRPFramebuffer) * n_frames)(*[f.ptr for f in fbs])
    uni = np.frombuffer(b"".join(d.uniform for d in draws), dtype=np.uint8).copy()
    stride = len(draws[0].uniform)
    prog.set_uniform(draws[0].uniform)
    lib.dll.srpB200SetSyncMode(H.SRP_B200_SYNC_EXPLICIT)
    call = lambda: lib.dll.srpB200DrawBatch(ib, vb, arr, n_frames, C.byref(prog.sp), uni.ctypes.data, stride,
                                            H.SRP_PRIM_TRIANGLES, 0, len(mesh[1]), 1)
    call(); lib.dll.srpB200Finish()
    lib.dll.srpB200ResetStats(); lib.dll.srpB200SetProfiling(1); lib.stage_times()
    stream = torch.cuda.ExternalStream(lib.dll.srpB200Stream())
    ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(steps)]
    with torch.cuda.stream(stream):
        for a, b in ev:
            a.record(stream); call(); b.record(stream)
    torch.cuda.synchronize()
    ms = sum(a.elapsed_time(b) for a, b in ev) / steps
    st = lib.stage_times(); lib.dll.srpB200SetProfiling(0)
    stats = lib.stats()
    lib.dll.srpB200SetSyncMode(H.SRP_B200_SYNC_DRAW)
    for f in fbs:
        f.free()
    return {"scene": f"cfg5 batch: {n_frames} teapot frames {size}x{size} in one srpB200DrawBatch ({mesh[2]})",
            "ms_per_batch": ms, "frames_per_s": n_frames / ms * 1e3, "stage_ms_per_batch": {k: st[k] / steps for k in ("geometry_ms", "binning_ms", "tiles_ms")},
            "frags_emitted_per_frame": stats["fragsEmitted"] / steps / n_frames, "gfrag_per_s": stats["fragsEmitted"] / steps / ms / 1e6,
            "framebuffer_bytes_per_frame": size * size * 9, "hbm_gbs_framebuffer_only": n_frames * size * size * 9 / ms / 1e6}


def main():
    lib = H.load_product()
    out = {"version": lib.dll.srpB200Version().decode(), "results": []}
    makers = {"cfg1": S.cfg1_textured_cube, "cfg2": S.cfg2_teapot, "cfg2_4k": lambda: S.cfg2_teapot(3840, 2160),
              "cfg3": S.cfg3_shell, "cfg3_r12": lambda: S.cfg3_shell(radius=1.2), "cfg4": S.cfg4_subpixel,
              "cfg5_frame": lambda: S.cfg5_frame(0)}
    only = sys.argv[1:]      # optional: names of the configs to run
    for name, make in makers.items():
        if only and name not in only:
            continue
        scene = make()
        r = time_scene(lib, scene)
        if name != "cfg4":      # (cfg4's worst-case scratch pools are 36 GB per lane)
            r["in_flight"] = time_in_flight(lib, scene)
        print(json.dumps(r)); out["results"].append(r)
    if not only or "cfg5" in only:
        r = time_batch(lib)
        print(json.dumps(r)); out["results"].append(r)
    Path(ROOT / "gpurun_out").mkdir(exist_ok=True)
    tag = ("_" + "_".join(only)) if only else ""
    (ROOT / "gpurun_out" / f"bench_all{tag}.json").write_text(json.dumps(out, indent=1))


if __name__ == "__main__":
    main()
