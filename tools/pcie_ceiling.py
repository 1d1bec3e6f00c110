#!/usr/bin/env python3
"""Host-link ceiling for the end-to-end figure (run under gpurun, alone or under torchrun):
pinned host <-> device copies of one cfg3 frame's traffic (28 MB up, 66 MB down) on every rank
at once -- D2H alone, H2D alone, both directions together -- as GB/s per rank and aggregate,
plus the frames/s those rates allow.  Also prints the GPU/NUMA topology the ranks see.
Writes gpurun_out/pcie_ceiling_n<world>.json (rank 0)."""
import json
import os
import subprocess
import sys
import time
from pathlib import Path

import torch
import torch.distributed as dist

ROOT = Path(__file__).resolve().parent.parent
UP, DOWN = 28_117_188, 66_355_200          # bytes per cfg3 frame: mesh up, colour + depth down


def main():
    rank = int(os.environ.get("RANK", "0")); world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    bind = os.environ.get("SRP_B200_NUMA_BIND", "1") != "0"
    note = "unbound"
    if bind:
        sys.path.insert(0, str(ROOT))
        from srp_b200 import numa
        note = numa.bind_to_gpu(local)
    hu = torch.empty(UP, dtype=torch.uint8).pin_memory(); du = torch.empty(UP, dtype=torch.uint8, device="cuda")
    hd = torch.empty(DOWN, dtype=torch.uint8).pin_memory(); dd = torch.empty(DOWN, dtype=torch.uint8, device="cuda")
    s_up, s_down = torch.cuda.Stream(), torch.cuda.Stream()

    def run(up, down, iters=30):
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier(); torch.cuda.synchronize()
        t0 = time.perf_counter()
        for _ in range(iters):
            if up:
                with torch.cuda.stream(s_up):
                    du.copy_(hu, non_blocking=True)
            if down:
                with torch.cuda.stream(s_down):
                    hd.copy_(dd, non_blocking=True)
        torch.cuda.synchronize()
        dt = time.perf_counter() - t0
        if world > 1:
            t = torch.tensor([dt], dtype=torch.float64, device="cuda")
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            dt = float(t[0])
        return dt / iters

    for _ in range(2):
        run(True, True, 5)
    t_up, t_down, t_both = run(True, False), run(False, True), run(True, True)
    if rank == 0:
        topo = ""
        try:
            topo = subprocess.run(["nvidia-smi", "topo", "-m"], capture_output=True, text=True, timeout=20).stdout
        except Exception as e:      # noqa: BLE001
            topo = f"(nvidia-smi topo failed: {e})"
        out = {"world": world, "numa_bind": note, "bytes_up": UP, "bytes_down": DOWN,
               "h2d_alone_gbs_per_rank": UP / t_up / 1e9, "d2h_alone_gbs_per_rank": DOWN / t_down / 1e9,
               "both_gbs_per_rank": (UP + DOWN) / t_both / 1e9,
               "frames_per_s_ceiling_per_rank": 1.0 / t_both, "frames_per_s_ceiling_aggregate": world / t_both,
               "frames_per_s_ceiling_d2h_only_aggregate": world / t_down,
               "host_cpus": len(os.sched_getaffinity(0)), "topology": topo}
        print(json.dumps({k: v for k, v in out.items() if k != "topology"}))
        print(topo)
        (ROOT / "gpurun_out").mkdir(exist_ok=True)
        (ROOT / "gpurun_out" / f"pcie_ceiling_n{world}{'' if bind else '_unbound'}.json").write_text(json.dumps(out, indent=1))
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
