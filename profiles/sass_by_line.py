#!/usr/bin/env python
"""Joins an ncu report's per-SASS-instruction counters with nvdisasm line info and prints executed warp-instructions per source line.

    python profiles/sass_by_line.py --rep gpurun_out/x.ncu-rep --obj lumenrenderer_b200/csrc/build/lb_restir.o --kernel k_ris [--instance 0] [--top 40] [--inline]

The object must be the one the report was captured with (same SASS). With --inline the line is reported with its inlined-at chain collapsed
to the innermost file:line (default) — the point is to see which expression of a fused kernel the instructions belong to."""
import argparse, collections, csv, io, os, re, subprocess, tempfile


def sass_lines(obj, kernel):
    tmp = tempfile.mkdtemp()
    subprocess.run(["cuobjdump", "-xelf", "all", os.path.abspath(obj)], cwd=tmp, capture_output=True)
    cubin = [f for f in os.listdir(tmp) if f.endswith(".cubin")][0]
    txt = subprocess.run(["nvdisasm", "-g", "-c", os.path.join(tmp, cubin)], capture_output=True, text=True).stdout
    out, cur, on = [], None, False
    for line in txt.splitlines():
        m = re.match(r"\s*\.text\.(\S+):", line)
        if m:
            on = kernel in m.group(1); continue
        if not on:
            continue
        m = re.match(r'\s*//## File "([^"]+)", line (\d+)', line)
        if m:
            cur = (os.path.basename(m.group(1)), int(m.group(2))); continue
        if re.match(r"\s*/\*[0-9a-f]{4,}\*/\s+\S", line):
            out.append(cur)
    return out


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--rep", required=True); ap.add_argument("--obj", required=True); ap.add_argument("--kernel", required=True)
    ap.add_argument("--instance", type=int, default=0); ap.add_argument("--top", type=int, default=40)
    ap.add_argument("--by-samples", action="store_true"); ap.add_argument("--symbol", help="substring of the mangled function name in the object (defaults to --kernel); needed for template instances")
    a = ap.parse_args()
    raw = subprocess.run(["ncu", "-i", a.rep, "--page", "source", "--csv", "--kernel-name", f"regex:{a.kernel}", "--print-source", "sass"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(raw)))
    starts = [i for i, r in enumerate(rows) if r and r[0] == "Kernel Name"]
    st = starts[a.instance]; en = starts[a.instance + 1] if a.instance + 1 < len(starts) else len(rows)
    hdr, data = rows[st + 1], rows[st + 2:en]
    ia, it, isamp = hdr.index("Instructions Executed"), hdr.index("Avg. Predicated-On Threads Executed"), hdr.index("# Samples")
    lines = sass_lines(a.obj, a.symbol or a.kernel)
    if len(lines) != len(data):
        print(f"warning: {len(lines)} SASS instructions in the object vs {len(data)} in the report")
    agg = collections.defaultdict(lambda: [0, 0.0, 0])
    tot = sum(int(r[ia]) for r in data); stot = sum(int(r[isamp]) for r in data)
    for loc, r in zip(lines, data):
        e = agg[loc]; n = int(r[ia]); e[0] += n; e[1] += n * float(r[it]); e[2] += int(r[isamp])
    print(f"{a.kernel} instance {a.instance}: {tot / 1e6:.1f} M warp instructions, {stot} samples")
    src_cache = {}
    for loc, (n, thr, s) in sorted(agg.items(), key=lambda kv: -(kv[1][2] if a.by_samples else kv[1][0]))[: a.top]:
        text = ""
        if loc:
            path = os.path.join("lumenrenderer_b200/csrc", loc[0])
            if path not in src_cache and os.path.exists(path):
                src_cache[path] = open(path).read().splitlines()
            if path in src_cache and loc[1] - 1 < len(src_cache[path]):
                text = src_cache[path][loc[1] - 1].strip()[:110]
        print(f"{100 * n / tot:5.1f}% inst  {100 * s / max(stot, 1):5.1f}% samples  thr {thr / max(n, 1):4.1f}  {loc[0] if loc else '?'}:{loc[1] if loc else 0}  {text}")


if __name__ == "__main__":
    main()
