#!/usr/bin/env python
"""Per-source-line executed-instruction / stall-sample shares of one launch, straight from ncu's own source
correlation (`ncu --page source --print-source cuda,sass`; needs --import-source on at capture time).

    python tools/ncu_srclines.py <report.ncu-rep> [top] [launch index]
"""
import subprocess
import sys


def main():
    rep = sys.argv[1]
    top = int(sys.argv[2]) if len(sys.argv) > 2 else 50
    cmd = ["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "cuda,sass"]
    if len(sys.argv) > 3:
        cmd += ["--launch-skip", sys.argv[3], "--launch-count", "1"]
    out = subprocess.run(cmd, capture_output=True, text=True).stdout
    f, fn, H, per = None, None, None, []
    for ln in out.splitlines():
        if ln.startswith('"File Path"'):
            f = ln.split('","')[1].rstrip('"').split("/")[-1]
        elif ln.startswith('"Function Name"'):
            fn = ln.split('","', 1)[1].rstrip('"')
        elif ln.startswith('"Line No"'):
            H = ln.strip('"').split('","')
        elif ln[:2] != '""' and '","-","-","' in ln:  # a source-line row (the SASS rows have an empty line number)
            head, tail = ln.split('","-","-","', 1)
            no, src = head.lstrip('"').split('","', 1)
            v = tail.rstrip('"').split('","')
            g = lambda name: float(v[H.index(name) - 4] or 0)
            per.append((f, int(no), src, g("Instructions Executed"), g("# Samples"), g("Thread Instructions Executed")))
    ti, ts, tt = (sum(p[k] for p in per) or 1.0 for k in (3, 4, 5))
    print("%s | warp instructions %.0f, thread instructions %.0f, samples %.0f" % (fn, ti, tt, ts))
    print("%-24s %7s %7s %8s  %s" % ("file:line", "inst %", "stall %", "thr/warp", "source"))
    for p in sorted(per, key=lambda p: -p[3])[:top]:
        print("%-24s %6.2f%% %6.2f%% %8.1f  %s" % ("%s:%d" % (p[0], p[1]), 100 * p[3] / ti, 100 * p[4] / ts, p[5] / max(p[3], 1), p[2].strip()[:110]))


if __name__ == "__main__":
    main()
