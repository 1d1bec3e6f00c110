"""CPU test of the strips' flag protocol (srp_b200/multigpu.py: StripTarget) with frames in flight.

The library calls are replaced by a recorder: every rank's calls go into one FIFO per (rank, lane)
-- a lane is a stream, so its operations run in order and different lanes are independent -- and
a small scheduler then "executes" the FIFOs against one shared array of flags, running any head
operation whose wait is satisfied.  Checked: nothing deadlocks; a rank writes into a ring slot
only after the root has consumed the slot's previous frame; the root consumes a frame only after
every rank has signalled it; flags of a slot only ever grow."""
import types

import pytest

from srp_b200 import multigpu as M


class FakeFramebuffer:
    def __init__(self, ident):
        self.ptr = ident

    def free(self):
        pass


class FakeDll:
    FLAGS = 0x10000

    def __init__(self, rank, ops):
        self.rank, self.ops, self.lane = rank, ops, 0

    def srpB200TileHeight(self): return 16
    def srpB200DeviceAlloc(self, nbytes): return self.FLAGS
    def srpB200DeviceFree(self, p): pass
    def srpB200FramebufferDevicePlane(self, fb, which): return 0x100000 + 16 * fb + which
    def srpB200IpcExport(self, ptr, buf): return 0
    def srpB200IpcOpen(self, buf): return self.FLAGS      # (every handle maps to "the" memory)
    def srpB200IpcClose(self, p): return 0
    def srpB200NewFramebufferOnDevice(self, w, h, c, d, s): return 1
    def srpB200SetLane(self, lane): self.lane = lane; return 0
    def srpB200GetLane(self): return self.lane
    def srpB200SetRowRange(self, a, b): pass
    def srpB200Finish(self): pass
    def _push(self, *op): self.ops.setdefault((self.rank, self.lane), []).append(op)
    def srpB200StreamSignal(self, ptr, value): self._push("signal", (ptr - self.FLAGS) // 4, value)
    def srpB200StreamWait(self, ptr, value): self._push("wait", (ptr - self.FLAGS) // 4, 1, value)
    def srpB200StreamWaitAll(self, ptr, count, value): self._push("wait", (ptr - self.FLAGS) // 4, count, value)


class FakeLib:
    def __init__(self, rank, ops):
        self.dll = FakeDll(rank, ops)
        self._n = 0

    def framebuffer(self, w, h):
        self._n += 1
        return FakeFramebuffer(self._n)


@pytest.mark.parametrize("world,ring,lanes,frames", [(2, 2, 1, 9), (2, 4, 2, 13), (4, 4, 4, 17), (8, 8, 4, 30)])
def test_strip_flags_protocol(monkeypatch, world, ring, lanes, frames):
    ops = {}
    box = {}
    current = {"rank": 0}
    fake_dist = types.SimpleNamespace(is_initialized=lambda: True, get_world_size=lambda group=None: world,
                                      get_rank=lambda group=None: current["rank"], barrier=lambda group=None: None)
    monkeypatch.setattr(M, "dist", fake_dist)
    monkeypatch.setattr("srp_b200.host.Framebuffer", lambda lib, ptr: FakeFramebuffer(ptr), raising=False)

    def exchange(b):      # the root's payload reaches everybody
        if b[0] is not None:
            box["payload"] = b[0]
        b[0] = box["payload"]

    targets = []
    for r in range(world):
        current["rank"] = r
        targets.append(M.StripTarget(FakeLib(r, ops), 640, 360, ring=ring, root=0, exchange=exchange, lanes=lanes))
    for k in range(frames):
        for r in range(world):
            current["rank"] = r
            t = targets[r]
            t.render(lambda fb, t=t, k=k: t.lib.dll._push("draw", k % ring, k))
            if r == 0:
                t.complete(consume=lambda fb, t=t, k=k: t.lib.dll._push("consume", k % ring, k))

    # "execute": a FIFO's head runs when its wait (if any) is satisfied
    flags = [0] * ((world + 1) * ring)
    consumed = {s: -1 for s in range(ring)}      # slot -> last frame the root has consumed from it
    pending = {q: list(v) for q, v in ops.items()}
    progress = True
    while progress and any(pending.values()):
        progress = False
        for q, fifo in pending.items():
            while fifo:
                op = fifo[0]
                if op[0] == "wait":
                    _, first, count, value = op
                    if not all(flags[first + i] >= value for i in range(count)):
                        break
                elif op[0] == "signal":
                    _, index, value = op
                    assert value > flags[index], "a slot's flag must only grow"
                    flags[index] = value
                elif op[0] == "draw":
                    _, slot, k = op
                    assert consumed[slot] == k - ring or k < ring, f"rank {q[0]} draws frame {k} into slot {slot} before frame {k - ring} was consumed"
                elif op[0] == "consume":
                    _, slot, k = op
                    base = slot * (world + 1)
                    assert all(flags[base + r] >= k + 1 for r in range(world)), f"frame {k} consumed before every rank signalled it"
                    consumed[slot] = k
                fifo.pop(0)
                progress = True
    assert not any(pending.values()), {q: v[:2] for q, v in pending.items() if v}
    for t in targets:
        t.free()
