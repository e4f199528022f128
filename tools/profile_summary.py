#!/usr/bin/env python
"""Turn one GPU visit (gpurun_out/<tag>/, written by tools/gpu_round.sh) into the tracked summary
profiles/<tag>.md: bench lines, the ncu launch list aggregated per kernel, the key metrics of the
`ncu --set full` capture of the dominant kernel and its hottest source lines.

    python tools/profile_summary.py <tag>
"""
import collections
import csv
import io
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))

RAW_KEYS = [
    "gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
    "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
    "smsp__issue_active.avg.pct_of_peak_sustained_active", "sm__warps_active.avg.pct_of_peak_sustained_active",
    "smsp__inst_executed.sum", "launch__registers_per_thread", "launch__grid_size", "launch__block_size",
    "launch__occupancy_limit_registers", "launch__occupancy_limit_shared_mem", "launch__waves_per_multiprocessor",
    "l1tex__t_bytes.sum", "lts__t_bytes.sum", "sm__cycles_elapsed.avg.per_second",
]


def launches_table(path):
    rows = list(csv.reader(l for l in open(path) if l.startswith('"')))
    h = rows[0]
    ki, vi, gi, bi = h.index("Kernel Name"), h.index("Metric Value"), h.index("Grid Size"), h.index("Block Size")
    agg = collections.OrderedDict()
    for r in rows[1:]:
        if len(r) <= vi:
            continue
        a = agg.setdefault(r[ki], [0, 0.0, r[gi], r[bi]])
        a[0] += 1
        a[1] += float(r[vi].replace(",", "")) / 1e3
    ours = {k: a for k, a in agg.items() if "k_" in k and "k_synth" not in k}
    tot = sum(a[1] for a in ours.values()) or 1.0
    out = ["| kernel | launches | total us | avg us | share of path | grid | block |", "|---|---|---|---|---|---|---|"]
    for k, a in agg.items():
        share = "%.1f %%" % (100 * a[1] / tot) if k in ours else "(setup)"
        out.append("| `%s` | %d | %.1f | %.1f | %s | %s | %s |" % (k.replace("<unnamed>::", "")[:60], a[0], a[1], a[1] / a[0],
                                                                 share, a[2], a[3]))
    return "\n".join(out)


def raw_table(rep):
    out = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(out)))
    if len(rows) < 3:
        return "(no raw page)"
    h, u = rows[0], rows[1]
    lines = []
    for r in rows[2:]:
        name = r[h.index("Kernel Name")].replace("<unnamed>::", "")
        lines.append("**%s** (launch id %s)\n" % (name, r[0]))
        lines.append("| metric | value | unit |\n|---|---|---|")
        for k in RAW_KEYS:
            if k in h:
                i = h.index(k)
                lines.append("| %s | %s | %s |" % (k, r[i], u[i]))
        lines.append("")
    return "\n".join(lines)


def traffic_of(rep, tag, pairs_per_launch=256):
    """{bench kernel name: dram bytes per launch} from the raw page.  The scan-2 pass of the incremental loop (k_pass2) is
    captured for the 7 iterations of one chunk (rebuilds and deltas differ a lot): their AVERAGE, like bench.py's
    avg_launch_ms; other kernels: first launch."""
    out = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(out)))
    if len(rows) < 3:
        return {}
    h, u = rows[0], rows[1]
    acc = {}
    for r in rows[2:]:
        name = r[h.index("Kernel Name")]
        key = None
        if "k_pass2<" in name or "k_pass<(bool)1" in name or "k_pass<true" in name or "k_pass<1," in name:
            key = "k_pass<scan2>"
        elif "k_pass<(bool)0" in name or "k_pass<false" in name or "k_pass<0," in name:
            key = "k_pass<scan1>"
        elif "k_loop" in name:
            key = "k_loop"
        if not key or (key in acc and key != "k_pass<scan2>"):
            continue

        def val(k):
            i = h.index(k)
            v = float(r[i].replace(",", ""))
            unit = u[i].lower()
            return v * {"byte": 1, "kbyte": 1e3, "mbyte": 1e6, "gbyte": 1e9}.get(unit, 1)
        try:
            acc.setdefault(key, []).append((val("dram__bytes_read.sum"), val("dram__bytes_write.sum")))
        except ValueError:
            pass
    res = {}
    for key, v in acc.items():
        res[key] = {"visit": tag, "pairs_per_launch": pairs_per_launch, "launches_averaged": len(v),
                    "dram_bytes_read": sum(a for a, _ in v) / len(v), "dram_bytes_write": sum(b for _, b in v) / len(v)}
    return res


def main():
    tag = sys.argv[1]
    d = os.path.join(ROOT, "gpurun_out", tag)
    os.makedirs(os.path.join(ROOT, "profiles"), exist_ok=True)
    md = ["# GPU visit `%s`" % tag, ""]
    for f in ("gpu.txt", "nproc.txt", "pytest_gpu.txt", "smoke.txt"):
        p = os.path.join(d, f)
        if os.path.exists(p):
            md += ["`%s`:" % f, "```", open(p).read().strip(), "```", ""]
    for f in ("bench.json", "bench_ref.json"):
        p = os.path.join(d, f)
        if os.path.exists(p) and os.path.getsize(p):
            md += ["## %s (CUDA-event timed, NOT under a profiler)" % f, "```json"]
            for line in open(p):
                line = line.strip()
                if line.startswith("{"):
                    md.append(json.dumps(json.loads(line), indent=1))
            md += ["```", ""]
    p = os.path.join(d, "launches.csv")
    if os.path.exists(p):
        md += ["## ncu launch list (`--metrics gpu__time_duration.sum --clock-control none`; cold-cache, serialised: "
               "compare shares)", "", launches_table(p), ""]
    reps = sorted((f for f in os.listdir(d) if f.endswith(".ncu-rep")), key=lambda f: (-len(f), f))  # prof.ncu-rep (7 iterations) last
    for rep in reps:
        md += ["## ncu --set full: %s" % rep, "", raw_table(os.path.join(d, rep)), ""]
        for launch in ((0, 6) if rep == "prof.ncu-rep" else (0,)):  # the scan-2 pass: a rebuild and the last delta launch
            r = subprocess.run([sys.executable, os.path.join(ROOT, "tools", "ncu_srclines.py"), os.path.join(d, rep), "32",
                                str(launch)], capture_output=True, text=True)
            md += ["Hottest source lines of launch %d (ncu's own source correlation: executed warp instructions / stall "
                   "samples / average active threads):" % launch, "```", "\n".join(l[:170] for l in r.stdout.strip().splitlines()), "```", ""]
    # raw launch list next to the summary; DRAM traffic of the dominant kernel for bench.py's roofline.traffic
    p = os.path.join(d, "launches.csv")
    if os.path.exists(p):
        import shutil
        shutil.copy(p, os.path.join(ROOT, "profiles", tag + "_launches.csv"))
    for rep in reps:
        t = traffic_of(os.path.join(d, rep), tag)
        if t:
            tp = os.path.join(ROOT, "profiles", "traffic.json")
            cur = json.load(open(tp)) if os.path.exists(tp) else {}
            cur.update(t)
            json.dump(cur, open(tp, "w"))
    out = os.path.join(ROOT, "profiles", tag + ".md")
    open(out, "w").write("\n".join(md) + "\n")
    print(out)


if __name__ == "__main__":
    main()
