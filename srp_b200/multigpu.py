"""Multi-GPU plumbing for the two partitionings of the draw path.

One process per GPU (torchrun), `torch.distributed` for the plumbing:

* frame-parallel -- independent frames; `frame_partition` hands every rank a contiguous
  block, there is no collective on the data path (buffers are broadcast once at load).
* sort-first strips -- one large frame; every rank runs the full geometry front-end (so
  primitive ids and barycentric chains are identical everywhere) and rasterises only the
  tile rows of its strip (`srpB200SetRowRange`); `gather_strips` collects the strips'
  planes on the root.  The only collective of the path.  Works on CUDA tensors over NCCL
  and on CPU tensors over gloo (the CPU tests).
"""
from __future__ import annotations

import torch
import torch.distributed as dist


def frame_partition(n_frames: int, world: int, rank: int) -> range:
    """contiguous, balanced: the first (n_frames % world) ranks get one frame more"""
    base, extra = divmod(n_frames, world)
    start = rank * base + min(rank, extra)
    return range(start, start + base + (1 if rank < extra else 0))


def strip_rows(height: int, tile_h: int, world: int, rank: int) -> tuple[int, int]:
    """[row0, row1) of rank's strip: whole tile rows, balanced, covering [0, height)"""
    tile_rows = (height + tile_h - 1) // tile_h
    part = frame_partition(tile_rows, world, rank)
    return min(part.start * tile_h, height), min(part.stop * tile_h, height)


def gather_strips(strip: torch.Tensor, height: int, tile_h: int, dst: int = 0, group=None):
    """`strip` = this rank's rows [row0, row1) of one plane, shape [rows, W].  Returns the
    full [height, W] plane on `dst` (None elsewhere).  Strips are padded to the tallest one
    so that one all_gather moves everything (equal-sized messages)."""
    world = dist.get_world_size(group)
    rank = dist.get_rank(group)
    rows = [strip_rows(height, tile_h, world, r) for r in range(world)]
    tallest = max(r1 - r0 for r0, r1 in rows)
    r0, r1 = rows[rank]
    assert strip.shape[0] == r1 - r0, (strip.shape, rows[rank])
    padded = strip.new_zeros((tallest,) + tuple(strip.shape[1:]))
    padded[: r1 - r0] = strip
    parts = [torch.empty_like(padded) for _ in range(world)]
    dist.all_gather(parts, padded, group=group)
    if rank != dst:
        return None
    return torch.cat([p[: b - a] for p, (a, b) in zip(parts, rows)], 0)


def gather_strips_inplace(planes, height: int, tile_h: int, dst: int = 0, group=None):
    """In-place form of the gather: `planes` = full-frame [height, W] tensors (e.g. zero-copy
    views of the framebuffer's device planes) of which this rank has rendered rows
    [row0, row1).  Every other rank sends its rows of every plane straight into the same rows
    of `dst`'s planes -- one batch of point-to-point transfers (ncclSend/ncclRecv over NVLink,
    or gloo on CPU), no staging copies, no padding.  Returns the list of in-flight requests'
    completion having been waited for; on `dst` the planes then hold the whole frame."""
    world = dist.get_world_size(group)
    rank = dist.get_rank(group)
    if world == 1:
        return planes
    rows = [strip_rows(height, tile_h, world, r) for r in range(world)]
    ops = []
    if rank == dst:
        for r, (a, b) in enumerate(rows):
            if r != dst and b > a:
                ops += [dist.P2POp(dist.irecv, p[a:b], r, group) for p in planes]
    else:
        a, b = rows[rank]
        if b > a:
            ops += [dist.P2POp(dist.isend, p[a:b], dst, group) for p in planes]
    if ops:
        for req in dist.batch_isend_irecv(ops):
            req.wait()
    return planes


def device_plane_tensor(lib, fb, which: int) -> torch.Tensor:
    """zero-copy torch view of a framebuffer's device plane (0 colour, 1 depth bits, 2 stencil)"""
    ptr = lib.dll.srpB200FramebufferDevicePlane(fb.ptr, which)
    dtype, itemsize = ((torch.int32, 4), (torch.int32, 4), (torch.uint8, 1))[which]
    n = fb.width * fb.height

    class _Iface:
        __cuda_array_interface__ = {"shape": (fb.height, fb.width), "typestr": "<i4" if itemsize == 4 else "|u1",
                                    "data": (int(ptr), False), "version": 3, "strides": None}
    t = torch.as_tensor(_Iface(), device="cuda")
    assert t.numel() == n and t.dtype == dtype
    return t
