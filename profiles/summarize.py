#!/usr/bin/env python
"""Turns an ncu report (gpurun_out/*.ncu-rep) and/or a launch list CSV into the markdown tables committed under profiles/.

    python profiles/summarize.py --rep gpurun_out/prof_a.ncu-rep --launches gpurun_out/launches_a.csv --out profiles/r01_x.md --title "..."
"""
import argparse
import collections
import csv
import io
import re
import subprocess
import os

METRICS = [("gpu__time_duration.sum", "time"), ("launch__registers_per_thread", "regs"), ("sm__warps_active.avg.pct_of_peak_sustained_active", "occ %"),
           ("smsp__issue_active.avg.pct_of_peak_sustained_active", "issue %"), ("smsp__warps_eligible.avg.per_cycle_active", "elig warps"),
           ("smsp__thread_inst_executed_per_inst_executed.ratio", "thr/inst"), ("smsp__inst_executed.sum", "warp inst"),
           ("dram__bytes_read.sum", "dram rd"), ("dram__bytes_write.sum", "dram wr"), ("gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "dram %"),
           ("lts__t_sector_hit_rate.pct", "L2 hit %"), ("l1tex__t_sector_hit_rate.pct", "L1 hit %")]


def short(name):
    m = re.search(r"(k_\w+)(<\d>)?", name)
    return (m.group(1) + (m.group(2) or "")) if m else name[:40]


def to_bytes(v, unit):
    scale = {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}.get(unit, 1)
    return float(v) * scale


def rep_table(path):
    raw = subprocess.run(["ncu", "-i", path, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(raw)))
    hdr, units, data = rows[0], rows[1], rows[2:]
    col = {h: i for i, h in enumerate(hdr)}
    out = ["| kernel | " + " | ".join(n for _, n in METRICS) + " |", "|---|" + "---|" * len(METRICS)]
    for r in data:
        cells = []
        for key, label in METRICS:
            if key not in col:
                cells.append("-"); continue
            v, u = r[col[key]].replace(",", ""), units[col[key]]
            if label == "time":
                ms = float(v) * {"ns": 1e-6, "us": 1e-3, "ms": 1.0, "s": 1e3}.get(u, 1.0)
                cells.append(f"{ms:.3f} ms")
            elif label.startswith("dram r") or label.startswith("dram w"):
                cells.append(f"{to_bytes(v, u) / 1e6:.1f} MB")
            elif label == "warp inst":
                cells.append(f"{float(v) / 1e6:.1f} M")
            else:
                cells.append(f"{float(v):.2f}" if "." in v else v)
        out.append(f"| {short(r[col['Kernel Name']])} | " + " | ".join(cells) + " |")
    return "\n".join(out)


def launches_per_frame(path):
    lines = [l for l in open(path) if not l.startswith("==")]
    names = [re.sub(r"<\d>", "", short(r["Kernel Name"])) for r in csv.DictReader(lines)]
    starts = [i for i, n in enumerate(names) if n == "k_raygen"]
    return collections.Counter(names[starts[-2]:starts[-1]]) if len(starts) >= 2 else None


def traffic_json(path, out, note, launches_csv=None):
    """dram__bytes_read.sum + dram__bytes_write.sum per kernel, summed over the launches of the captured frame (bench.py reads this for `traffic`)."""
    import json
    raw = subprocess.run(["ncu", "-i", path, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(raw)))
    hdr, units, data = rows[0], rows[1], rows[2:]
    col = {h: i for i, h in enumerate(hdr)}
    agg, launches = collections.OrderedDict(), collections.Counter()
    for r in data:
        k = re.sub(r"<\d>", "", short(r[col["Kernel Name"]]))
        b = sum(to_bytes(r[col[m]].replace(",", ""), units[col[m]]) for m in ("dram__bytes_read.sum", "dram__bytes_write.sum"))
        agg[k] = agg.get(k, 0.0) + b; launches[k] += 1
    per_frame = launches_per_frame(launches_csv) if launches_csv else None
    if per_frame:       # the capture window need not be aligned to a frame: average per launch, times the launches one frame makes
        agg = collections.OrderedDict((k, v / launches[k] * per_frame.get(k, launches[k])) for k, v in agg.items())
        launches = per_frame
    json.dump({"source": os.path.basename(path), "note": note, "launches_per_frame": dict(launches), "dram_bytes_per_frame": agg}, open(out, "w"), indent=1)
    print("wrote", out)


def launches_table(path):
    lines = [l for l in open(path) if not l.startswith("==")]
    rows = list(csv.DictReader(lines))
    names = [r["Kernel Name"] for r in rows]
    starts = [i for i, n in enumerate(names) if "k_raygen" in n]
    if len(starts) < 2:
        return "(fewer than two frames captured)"
    frame = rows[starts[-2]:starts[-1]]
    agg = collections.OrderedDict()
    for r in frame:
        v = float(r["Metric Value"].replace(",", "")); u = r["Metric Unit"]
        ms = v * {"ns": 1e-6, "nsecond": 1e-6, "us": 1e-3, "usecond": 1e-3, "ms": 1.0, "msecond": 1.0}.get(u, 1e-6)
        k = short(r["Kernel Name"]); a = agg.setdefault(k, [0, 0.0]); a[0] += 1; a[1] += ms
    total = sum(a[1] for a in agg.values())
    out = ["| kernel | launches/frame | ms/frame (ncu, cold-cache, serialised) | share |", "|---|---|---|---|"]
    for k, (n, ms) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
        out.append(f"| {k} | {n} | {ms:.3f} | {100 * ms / total:.1f} % |")
    out.append(f"| **total** | {sum(a[0] for a in agg.values())} | {total:.3f} | 100 % |")
    return "\n".join(out)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--rep"); ap.add_argument("--launches"); ap.add_argument("--out", required=True); ap.add_argument("--title", default="ncu summary"); ap.add_argument("--notes", default=""); ap.add_argument("--traffic-json")
    a = ap.parse_args()
    if a.traffic_json and a.rep:
        traffic_json(a.rep, a.traffic_json, a.title, a.launches)
    parts = [f"# {a.title}", ""]
    if a.notes:
        parts += [a.notes, ""]
    if a.launches:
        parts += ["## Launch list of one steady-state frame (`ncu --metrics gpu__time_duration.sum --clock-control none`)", "", launches_table(a.launches), ""]
    if a.rep:
        parts += ["## `ncu --set full --clock-control none` per launch", "", rep_table(a.rep), ""]
    open(a.out, "w").write("\n".join(parts))
    print("wrote", a.out)


if __name__ == "__main__":
    main()
