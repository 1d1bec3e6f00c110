#!/usr/bin/env python3
"""profiles/r02_multigpu_summary.md from the bench lines and host-link ceilings kept under profiles/."""
import json
import sys
from pathlib import Path

P = Path(__file__).resolve().parent.parent / "profiles"


def load(p):
    return json.loads((P / p).read_text().strip().splitlines()[-1])


def main():
    tag = sys.argv[1] if len(sys.argv) > 1 else "r02"
    ns = [n for n in (2, 4, 8) if (P / f"{tag}_bench_cfg3_n{n}.json").exists()]
    one = load(f"{tag}_bench_cfg3_n1.json") if (P / f"{tag}_bench_cfg3_n1.json").exists() else None
    L = [f"# Round 2 -- multi-GPU evidence (builder's own runs: `gpurun --gpus N -- tools/gpu_multi.sh N {tag}`)", "",
         "Bench lines: " + ", ".join(f"`{tag}_bench_cfg3_n{n}.json`" for n in ([1] if one else []) + ns)
         + " (torchrun, one rank per GPU, NCCL saw N ranks).  All device times are the max over ranks.", "",
         "## Frame-parallel (headline `value`, weak scaling) and BASELINE config 5", "",
         "| N | cfg3 frames/s (kernel-only, 4 frames in flight per GPU) | per GPU | cfg5: 1024 teapot frames 1024x1024, frames/s | per-GPU algorithmic GB/s (cfg5) | batch == single draw |",
         "|---|---|---|---|---|---|"]
    rows = ([(1, one)] if one else []) + [(n, load(f"{tag}_bench_cfg3_n{n}.json")) for n in ns]
    for n, d in rows:
        c5 = d.get("secondary", {}).get("cfg5_batch_frame_parallel", {})
        if "frames_per_s" in c5:
            L.append(f"| {n} | {d['value']:.0f} | {d['value'] / n:.0f} | {c5['frames_per_s']:.0f} ({c5['ms_per_batch']:.3f} ms for the 1024 frames) | "
                     f"{c5['hbm_gbs_per_gpu_algorithmic']:.0f} | {c5['batch_frame_equals_single_draw']} |")
        else:
            L.append(f"| {n} | {d['value']:.0f} | {d['value'] / n:.0f} | {c5} | - | - |")
    L += ["", "## Sort-first strips: ONE cfg3 frame split over N GPUs", "",
          "| N | fused: strips written into the root's planes over NVLink (CUDA IPC peer memory, flags in root memory); speed-up against one GPU rendering one frame at a time | NCCL send/recv gather (baseline) | bit-exact vs one GPU | NVLink bytes/frame (fused) |",
          "|---|---|---|---|---|"]
    for n, d in rows:
        if n == 1 or "strips" not in d or "fused_peer_write" not in d["strips"]:
            continue
        f, g = d["strips"]["fused_peer_write"], d["strips"]["nccl_gather"]
        single = d.get("one_frame_in_flight", {}).get("value", d["value"]) / n      # one GPU rendering one frame at a time
        flight = f.get("frames_in_flight", 1)
        L.append(f"| {n} | {f['frames_per_s']:.0f} frames/s ({f['ms_per_frame']:.3f} ms; {flight} frame(s) in flight) = {f['frames_per_s'] / single:.2f}x one GPU one frame at a time"
                 + (f", {f['frames_per_s'] / (d['value'] / n):.2f}x one GPU with 4 frames in flight" if flight > 1 else "") + " | "
                 f"{g['frames_per_s']:.0f} frames/s ({g['ms_per_frame']:.3f} ms) | {f['bit_exact_vs_single_gpu']} / {g['bit_exact_vs_single_gpu']} | {f['nvlink_bytes_per_frame'] / 1e6:.1f} MB |")
    L += ["", "The geometry front-end (~0.12 ms one frame at a time) is replicated on every rank by construction (identical primitive ids and "
          "barycentric chains), so it bounds the strips' speed-up; the tile kernel's share shrinks with N and the peer write-back costs no extra pass.  "
          "With frames in flight (StripTarget(lanes=4): frame k on lane k % 4, per-slot flags) the front-end of one frame runs under the tiles of another.", "",
          "## End to end (host buffers in, host-visible planes out, every step's copies timed) against the host link's ceiling", "",
          f"`tools/pcie_ceiling.py` copies one frame's traffic (28 MB up, 66 MB down, pinned) on every rank at once: `{tag}_pcie_ceiling_n*.json`.  "
          f"The box is a KVM guest with ONE virtual NUMA node (`{tag}_topology_n8.txt`: every GPU reports the same CPU affinity / NUMA 0), so there is "
          "no placement to fix from inside the guest (`srp_b200/numa.py` binds where a box does expose the topology): the aggregate host-link bandwidth is what saturates.", "",
          "| N | D2H alone, GB/s per rank | both directions, GB/s per rank | ceiling, frames/s (aggregate) | measured e2e, frames/s (two frames in flight) | fraction of ceiling |",
          "|---|---|---|---|---|---|"]
    for n, d in rows:
        cp = P / f"{tag}_pcie_ceiling_n{n}.json"
        if not cp.exists():
            continue
        c = json.loads(cp.read_text())
        L.append(f"| {n} | {c['d2h_alone_gbs_per_rank']:.1f} | {c['both_gbs_per_rank']:.1f} | {c['frames_per_s_ceiling_aggregate']:.0f} | "
                 f"{d['e2e']['value']:.0f} | {d['e2e']['value'] / c['frames_per_s_ceiling_aggregate']:.2f} |")
    L += ["", "The end-to-end figure sits at the measured ceiling of the host link at every N (ceiling and bench are separate runs on separately "
          "leased boxes, hence fractions around 1); per-rank D2H bandwidth falls from ~56 GB/s (N = 1) to ~12 GB/s (N = 8) because the eight x16 "
          "links share the guest's host-memory path.  The library's part -- uploads on their own stream behind the last use of the buffer, "
          "downloads on a copy stream, two frames in flight -- is not the limiter.", ""]
    (P / f"{tag}_multigpu_summary.md").write_text("\n".join(L))
    print("wrote", P / f"{tag}_multigpu_summary.md")


if __name__ == "__main__":
    main()
