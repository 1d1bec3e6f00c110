"""GPU test of the multi-process strips path: two processes -- both on cuda:0, the
mechanism is the same as between two GPUs -- render the two strips of one frame; the non-root
process maps the root's framebuffer planes through CUDA IPC and its tile kernel writes its strip
straight into the root's memory; flags in the root's memory signal completion and release the
ring slots.  The assembled frames must equal the single-process render bit for bit."""
import os
import socket

import numpy as np
import pytest
import torch.distributed as dist
import torch.multiprocessing as mp

pytestmark = pytest.mark.gpu


def _worker(rank, world, port, out, ring, lanes):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), SRP_B200_DEVICE="0")
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        from srp_b200 import host as H, multigpu as M, scenes as S
        lib = H.load_product()
        frames = [S.cfg3_shell(640, 360, n=96, radius=r) for r in (3.0, 1.4, 2.2, 1.2, 2.6, 1.7, 2.9, 1.3, 2.0)]
        want = [S.render(lib, sc) for sc in frames] if rank == 0 else None
        lib.dll.srpB200SetSyncMode(H.SRP_B200_SYNC_EXPLICIT)
        target = M.StripTarget(lib, 640, 360, ring=ring, root=0, lanes=lanes)
        preps = [S.Prepared(lib, sc) for sc in frames]
        got = []
        for k, p in enumerate(preps):
            def draw(fb, p=p):
                p.fb, keep = fb, p.fb
                try:
                    p.draw_all()
                finally:
                    p.fb = keep
            fb = target.render(draw)
            if rank == 0:
                # the consumer of the complete frame: bring it to the host (stream-ordered), then release the slot
                def consume(f):
                    lib.dll.srpB200FramebufferDownloadAsync(f.ptr)
                    lib.dll.srpB200FramebufferFence(f.ptr)      # the slot is released when the frame has left it
                target.complete(consume=consume)
                lib.dll.srpB200FramebufferWait(fb.ptr)
                got.append(fb.planes(download=False))
        lib.dll.srpB200Finish()
        ok = True
        if rank == 0:
            for k in range(len(frames)):
                for a, b in zip(got[k], want[k]):
                    ok = ok and bool(np.array_equal(a, b))
            ok = ok and lib.stats()["overflow"] == 0
        for p in preps:
            p.free()
        target.free()
        lib.dll.srpB200SetSyncMode(H.SRP_B200_SYNC_DRAW)
        out[rank] = ok
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("ring,lanes", [(2, 1), (4, 2)], ids=["one_lane", "two_frames_in_flight"])
def test_strips_written_into_the_roots_framebuffer_over_ipc(ring, lanes):
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        port = s.getsockname()[1]
    manager = mp.Manager()
    out = manager.dict()
    mp.spawn(_worker, args=(2, port, out, ring, lanes), nprocs=2, join=True)
    assert dict(out) == {0: True, 1: True}
