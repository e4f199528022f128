#!/usr/bin/env python
"""Per-source-line instruction / stall-sample shares of one kernel, from an ncu report.

    python tools/ncu_lines.py <report.ncu-rep> <kernel-regex> [top]

Joins `ncu --page source --csv` (SASS rows with executed-instruction and sampling counts) with the line table
of the in-tree library (`nvdisasm -g`), so it works without a GPU.  The library must be the build that was
profiled (compile with -lineinfo)."""
import collections
import csv
import io
import os
import re
import subprocess
import sys
import tempfile

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIB = os.path.join(ROOT, "icet_b200", "lib", "libicet_b200.so")


def line_table(kernel_re):
    tmp = tempfile.mkdtemp()
    subprocess.run(["cuobjdump", "-xelf", "all", LIB], cwd=tmp, check=True, stdout=subprocess.DEVNULL)
    cubin = [f for f in os.listdir(tmp) if f.endswith(".cubin")][0]
    sass = subprocess.run(["nvdisasm", "-g", "-c", os.path.join(tmp, cubin)], capture_output=True, text=True).stdout
    table, cur, loc, active = {}, None, None, False
    for ln in sass.splitlines():
        m = re.match(r"\s*\.section\s+\.text\.(\S+?),", ln)
        if m:
            name = subprocess.run(["c++filt", m.group(1)], capture_output=True, text=True).stdout.strip()
            active = re.search(kernel_re, name) is not None and cur is None
            if active:
                cur = name
            continue
        if not active:
            continue
        m = re.match(r'\s*//## File "([^"]+)", line (\d+)', ln)
        if m:
            loc = (os.path.basename(m.group(1)), int(m.group(2)))
            continue
        m = re.match(r"\s*/\*([0-9a-f]{4,})\*/\s+(.*);", ln)
        if m:
            table[int(m.group(1), 16)] = (loc, m.group(2).strip())
    return cur, table


def main():
    rep, kre = sys.argv[1], sys.argv[2]
    top = int(sys.argv[3]) if len(sys.argv) > 3 else 40
    out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(out)))
    ker, H, data = None, None, collections.OrderedDict()
    for r in rows:
        if not r:
            continue
        if r[0] == "Kernel Name":
            ker = r[1]
            data.setdefault(ker, [])
            continue
        if r[0] == "Address":
            H = r
            continue
        if ker:
            data[ker].append(r)
    ker = next(k for k in data if re.search(kre, k))
    R = data[ker]
    ix, ism = H.index("Instructions Executed"), H.index("# Samples")
    base = min(int(r[0], 16) for r in R)
    name, table = line_table(kre)
    per = collections.defaultdict(lambda: [0.0, 0.0])
    files = {}
    for r in R:
        off = int(r[0], 16) - base
        loc = table.get(off, (("?", 0), ""))[0] or ("?", 0)
        per[loc][0] += float(r[ix] or 0)
        per[loc][1] += float(r[ism] or 0)
    ti = sum(v[0] for v in per.values())
    ts = sum(v[1] for v in per.values()) or 1.0
    print("kernel:", ker, "| SASS instructions:", len(R))
    print("%-22s %8s %8s  %s" % ("file:line", "inst %", "stall %", "source"))
    for loc, v in sorted(per.items(), key=lambda kv: -kv[1][0])[:top]:
        f, l = loc
        if f not in files:
            p = [os.path.join(ROOT, "icet_b200", "csrc", f), os.path.join(ROOT, "include", f)]
            p = [q for q in p if os.path.exists(q)]
            files[f] = open(p[0]).read().splitlines() if p else []
        src = files[f][l - 1].strip() if 0 < l <= len(files[f]) else ""
        print("%-22s %7.2f%% %7.2f%%  %s" % ("%s:%d" % (f, l), 100 * v[0] / ti, 100 * v[1] / ts, src[:110]))


if __name__ == "__main__":
    main()
