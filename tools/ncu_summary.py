#!/usr/bin/env python3
"""Turn an Nsight Compute report (+ the launch-list CSV of the same command) into the
markdown summary kept under profiles/.   usage: ncu_summary.py REPORT.ncu-rep LAUNCHES.csv OUT.md [title]"""
import collections
import csv
import io
import subprocess
import sys

METRICS = [
    ("gpu__time_duration.sum", "duration"),
    ("dram__bytes_read.sum", "DRAM read"),
    ("dram__bytes_write.sum", "DRAM write"),
    ("gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "DRAM % of peak"),
    ("sm__throughput.avg.pct_of_peak_sustained_elapsed", "SM % of peak"),
    ("sm__warps_active.avg.pct_of_peak_sustained_active", "warps active % (occupancy)"),
    ("launch__registers_per_thread", "registers/thread"),
    ("launch__grid_size", "grid"),
    ("launch__block_size", "block"),
    ("smsp__inst_executed.sum", "warp instructions"),
    ("smsp__thread_inst_executed_per_inst_executed.ratio", "active threads / instruction"),
    ("l1tex__t_bytes.sum", "L1 bytes"),
    ("lts__t_bytes.sum", "L2 bytes"),
]


def ncu(*args):
    return subprocess.run(["ncu", *args], capture_output=True, text=True).stdout


def main():
    rep, launches, out = sys.argv[1:4]
    title = sys.argv[4] if len(sys.argv) > 4 else rep
    lines = [f"# {title}", ""]
    # ---- launch list
    rows = [l for l in open(launches) if not l.startswith("==")] if launches != "/dev/null" else []
    agg = collections.OrderedDict()
    for r in csv.DictReader(rows):
        agg.setdefault(r["Kernel Name"].split("(")[0], []).append(float(r["Metric Value"].replace(",", "")))
    total = sum(sum(v) for k, v in agg.items() if "srpd" in k)
    if rows:
        lines += ["## Launch list (`ncu --metrics gpu__time_duration.sum --clock-control none`; cold-cache, serialised: compare shares)", "",
                  "| kernel | launches | avg us | share of srpd* time |", "|---|---|---|---|"]
    for k, v in agg.items():
        share = f"{100 * sum(v) / total:.1f} %" if "srpd" in k else "-"
        lines.append(f"| `{k}` | {len(v)} | {sum(v) / len(v) / 1e3:.1f} | {share} |")
    # ---- per-kernel metrics
    raw = list(csv.reader(io.StringIO(ncu("-i", rep, "--page", "raw", "--csv"))))
    hdr, units = raw[0], raw[1]
    idx = {h: i for i, h in enumerate(hdr)}
    lines += ["", "## `ncu --set full` of one frame's kernels", ""]
    for r in raw[2:]:
        lines.append(f"### `{r[idx['Kernel Name']].split('(')[0]}`")
        lines += ["", "| metric | value |", "|---|---|"]
        for m, label in METRICS:
            if m in idx:
                lines.append(f"| {label} (`{m}`) | {r[idx[m]]} {units[idx[m]]} |")
        lines.append("")
    # ---- hottest source lines of the two big kernels
    for kern in ("srpdTileKernel", "srpdGeomKernel"):
        src = list(csv.reader(io.StringIO(ncu("-i", rep, "--page", "source", "--print-source", "cuda,sass", "--csv", "--kernel-name", f"regex:{kern}"))))
        cur, items = None, []
        for r in src:
            if len(r) >= 2 and r[0] == "File Path":
                cur = r[1].split("/")[-1]
            elif len(r) >= 8 and r[0].isdigit():
                try:
                    items.append((int(r[7]), int(r[4]) if r[4].isdigit() else 0, cur, int(r[0]), r[1].strip()[:100]))
                except ValueError:
                    pass
        if not items:
            continue
        ti, ts = sum(i[0] for i in items), max(1, sum(i[1] for i in items))
        lines += [f"### hottest source lines of `{kern}` (share of executed warp instructions / of stall samples)", "", "```"]
        for i in sorted(items, reverse=True)[:18]:
            lines.append(f"{100 * i[0] / ti:5.1f}% inst {100 * i[1] / ts:5.1f}% stall  {i[2]}:{i[3]}  {i[4]}")
        lines += ["```", ""]
    # per-kernel DRAM traffic of one launch, for bench.py's roofline.traffic
    import json
    def to_bytes(v, unit):
        f = float(v.replace(",", ""))
        return f * {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}.get(unit, 1)
    traffic = {}
    for r in raw[2:]:
        name = r[idx["Kernel Name"]].split("(")[0].replace("void ", "").split("<")[0]
        if name in traffic:
            continue
        rd = to_bytes(r[idx["dram__bytes_read.sum"]], units[idx["dram__bytes_read.sum"]])
        wr = to_bytes(r[idx["dram__bytes_write.sum"]], units[idx["dram__bytes_write.sum"]])
        traffic[name] = {"dram_read_bytes": rd, "dram_write_bytes": wr, "dram_bytes": rd + wr,
                         "duration_us": float(r[idx["gpu__time_duration.sum"]].replace(",", ""))}
    json.dump({"source": rep.split("/")[-1], "note": "one launch each, ncu --set full --clock-control none (cold cache, serialised)",
               "kernels": traffic}, open(out.replace(".md", "_traffic.json"), "w"), indent=1)
    open(out, "w").write("\n".join(lines) + "\n")
    print("wrote", out)


if __name__ == "__main__":
    main()
