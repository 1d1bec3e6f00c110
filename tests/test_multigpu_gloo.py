"""CPU tests (world_size 2, gloo) of the multi-GPU host logic: frame partition, strip rows,
and the strip gather -- the only collective of the path."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from srp_b200 import multigpu as M


def test_frame_partition_covers_everything_once():
    for n in (0, 1, 7, 1024, 1025):
        for world in (1, 2, 3, 8):
            parts = [M.frame_partition(n, world, r) for r in range(world)]
            flat = [f for p in parts for f in p]
            assert flat == list(range(n))
            assert max(len(p) for p in parts) - min(len(p) for p in parts) <= 1


def test_strip_rows_are_tile_aligned_and_cover_the_frame():
    for height in (16, 100, 360, 1080, 2160):
        for world in (1, 2, 4, 8):
            rows = [M.strip_rows(height, 16, world, r) for r in range(world)]
            assert rows[0][0] == 0 and rows[-1][1] == height
            for (a0, a1), (b0, b1) in zip(rows[:-1], rows[1:]):
                assert a1 == b0
            assert all(r0 % 16 == 0 for r0, r1 in rows if r1 > r0), "non-empty strips start on a tile row"


def _worker(rank, world, port, height, width, out):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        full = (torch.arange(height * width, dtype=torch.int32).reshape(height, width) * 2654435761 % 1000003).to(torch.int32)
        r0, r1 = M.strip_rows(height, 16, world, rank)
        mine = full[r0:r1].clone()                       # what this rank "rendered"
        got = M.gather_strips(mine, height, 16, dst=0)
        # in-place form: every rank holds full-size planes of which only its own rows are valid
        planes = []
        for k in range(3):
            pl = torch.full((height, width), -7, dtype=torch.int32)
            pl[r0:r1] = full[r0:r1] + k
            planes.append(pl)
        M.gather_strips_inplace(planes, height, 16, dst=0)
        inplace_ok = rank != 0 or all(bool(torch.equal(pl, full + k)) for k, pl in enumerate(planes))
        # frame-parallel bookkeeping: every rank reports which frames it owns
        frames = torch.zeros(37, dtype=torch.int32)
        frames[list(M.frame_partition(37, world, rank))] = 1
        dist.all_reduce(frames)
        ok = bool((frames == 1).all()) and inplace_ok
        if rank == 0:
            ok = ok and got is not None and bool(torch.equal(got, full))
        else:
            ok = ok and got is None
        out[rank] = ok
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("height", [360, 100])
def test_strip_gather_world2_gloo(height):
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        port = s.getsockname()[1]
    manager = mp.Manager()
    out = manager.dict()
    mp.spawn(_worker, args=(2, port, height, 64, out), nprocs=2, join=True)
    assert dict(out) == {0: True, 1: True}
