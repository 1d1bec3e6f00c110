"""Multi-GPU plumbing for the two partitionings of the draw path.

One process per GPU (torchrun), `torch.distributed` for the plumbing:

* frame-parallel -- independent frames; `frame_partition` hands every rank a contiguous
  block, there is no collective on the data path (buffers are broadcast once at load).
* sort-first strips -- one large frame; every rank runs the full geometry front-end (so
  primitive ids and barycentric chains are identical everywhere; primitives whose box misses
  the rank's rows keep their id but store nothing) and rasterises only the tile rows of its
  strip (`srpB200SetRowRange`).  Two ways to assemble the frame on the root:
    - `StripTarget` (the product path): the framebuffer lives on the root GPU, the other ranks
      map its planes through CUDA IPC and their tile kernels write their strips straight into
      the root's memory over NVLink as part of the tile write-back; completion and
      back-pressure go through 32-bit flags in the root's memory (srpB200StreamSignal /
      srpB200StreamWait).  No staging copy, no separate gather step, no host round trip.
    - `gather_strips` / `gather_strips_inplace` (the NCCL baseline named by `north_star`):
      every rank renders into its own framebuffer and one batch of ncclSend / ncclRecv moves
      the strips to the root.  Works on CUDA tensors over NCCL and on CPU tensors over gloo
      (the CPU tests).
"""
from __future__ import annotations

import ctypes as C

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


class StripTarget:
    """A ring of `ring` root-owned framebuffers that every rank renders its strip into.

    Root: creates the framebuffers and a block of 32-bit flags per ring slot -- flags[slot][r] =
    frames rank r has finished writing into the slot (as frame number + 1), flags[slot][world] =
    frames the root has consumed from it -- and exports the three planes of every framebuffer and
    the flags as CUDA IPC handles (srpB200IpcExport); `exchange` is any function that hands the
    root's bytes to all ranks (a torch.distributed broadcast of objects, over NCCL or gloo).
    Other ranks: open the handles and wrap the mapped planes in framebuffers of their own
    (srpB200NewFramebufferOnDevice): the tile kernel's write-back then stores into the root's
    memory.

    Frame k (every rank), on lane k % lanes and in slot k % ring (`ring` is a multiple of `lanes`,
    so a slot always lives on one lane and its flags only ever grow): wait until the root has
    consumed the slot's previous frame, srpB200SetRowRange(own rows), clear + draw into the slot,
    signal flags[slot][rank] = k + 1.  Root, additionally: wait for flags[slot][r] >= k + 1 of
    every rank -- the frame is complete in its memory --, run `consume` (e.g. an asynchronous
    download), signal flags[slot][world] = k + 1.  With lanes > 1 consecutive frames are in flight
    side by side: the geometry front-end, which every rank runs in full, overlaps the previous
    frame's tiles.  Everything is enqueued on the lanes' streams; nothing blocks the host."""

    def __init__(self, lib, width, height, ring=2, root=0, group=None, exchange=None, lanes=1):
        if lanes < 1 or ring % lanes:
            raise ValueError("ring must be a multiple of lanes")
        self.lib, self.width, self.height, self.ring, self.root, self.group = lib, width, height, ring, root, group
        self.lanes = lanes
        self.world = dist.get_world_size(group) if dist.is_initialized() else 1
        self.rank = dist.get_rank(group) if dist.is_initialized() else 0
        self.tile_h = int(lib.dll.srpB200TileHeight())
        self.rows = strip_rows(height, self.tile_h, self.world, self.rank)
        self._mapped = []
        d = lib.dll
        payload = None
        if self.rank == root:
            self.fbs = [lib.framebuffer(width, height) for _ in range(ring)]
            self.flags = d.srpB200DeviceAlloc(4 * (self.world + 1) * ring)
            if not self.flags:
                raise RuntimeError("srpB200DeviceAlloc failed")
            handles = []
            for fb in self.fbs:
                for which in range(3):
                    handles.append(self._export(d.srpB200FramebufferDevicePlane(fb.ptr, which)))
            payload = {"planes": handles, "flags": self._export(self.flags)}
        if self.world > 1:
            box = [payload]
            (exchange or self._broadcast)(box)
            payload = box[0]
        if self.rank != root:
            planes = [self._open(h) for h in payload["planes"]]
            self.flags = self._open(payload["flags"])
            self.fbs = []
            for i in range(ring):
                ptr = d.srpB200NewFramebufferOnDevice(width, height, planes[3 * i], planes[3 * i + 1], planes[3 * i + 2])
                if not ptr:
                    raise RuntimeError("srpB200NewFramebufferOnDevice failed")
                from .host import Framebuffer
                self.fbs.append(Framebuffer(lib, ptr))
        self.frame = 0

    def _broadcast(self, box):
        dist.broadcast_object_list(box, src=self.root, group=self.group)

    def _export(self, device_ptr) -> bytes:
        buf = (C.c_ubyte * 64)()
        if self.lib.dll.srpB200IpcExport(device_ptr, buf) != 0:
            raise RuntimeError("srpB200IpcExport failed")
        return bytes(buf)

    def _open(self, handle: bytes):
        buf = (C.c_ubyte * 64).from_buffer_copy(handle)
        p = self.lib.dll.srpB200IpcOpen(buf)
        if not p:
            raise RuntimeError("srpB200IpcOpen failed: peer access between the GPUs is required")
        self._mapped.append(p)
        return p

    def flag_ptr(self, slot: int, index: int) -> int:
        return int(self.flags) + 4 * (slot * (self.world + 1) + index)

    def render(self, draw):
        """enqueue one frame: `draw(fb)` issues srpFramebufferClear + the frame's draws into fb.
        Returns the framebuffer that will hold the complete frame on the root."""
        d = self.lib.dll
        k = self.frame
        slot = k % self.ring
        fb = self.fbs[slot]
        keep = d.srpB200GetLane()
        d.srpB200SetLane(k % self.lanes)
        try:
            if k >= self.ring:
                d.srpB200StreamWait(self.flag_ptr(slot, self.world), k - self.ring + 1)      # the slot's previous frame was consumed
            d.srpB200SetRowRange(self.rows[0], self.rows[1])
            try:
                draw(fb)
            finally:
                d.srpB200SetRowRange(0, 2 ** 64 - 1)
            d.srpB200StreamSignal(self.flag_ptr(slot, self.rank), k + 1)
        finally:
            d.srpB200SetLane(keep)
        self.frame = k + 1
        return fb

    def complete(self, consume=None):
        """root only: order the frame's lane behind every rank's strip of the frame enqueued last,
        run `consume(fb)` (enqueue-only work on the complete frame), release the slot"""
        assert self.rank == self.root
        d = self.lib.dll
        k = self.frame - 1
        slot = k % self.ring
        keep = d.srpB200GetLane()
        d.srpB200SetLane(k % self.lanes)
        try:
            # every rank's flag of the slot in one launch (the root's own was signalled on this lane just before)
            d.srpB200StreamWaitAll(self.flag_ptr(slot, 0), self.world, k + 1)
            if consume is not None:
                consume(self.fbs[slot])
            d.srpB200StreamSignal(self.flag_ptr(slot, self.world), k + 1)
        finally:
            d.srpB200SetLane(keep)

    def free(self):
        d = self.lib.dll
        d.srpB200Finish()
        if dist.is_initialized() and self.world > 1:
            dist.barrier(group=self.group)      # nobody unmaps or frees while a peer may still write
        for fb in self.fbs:
            fb.free()
        if self.rank == self.root:
            d.srpB200DeviceFree(self.flags)
        else:
            for p in self._mapped:
                d.srpB200IpcClose(p)
        self.fbs, self._mapped = [], []
