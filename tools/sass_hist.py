#!/usr/bin/env python
"""SASS instruction histogram of the hot kernels of icet_b200/lib/libicet_b200.so (cuobjdump -sass; no GPU needed), so that
instruction-count claims can be checked without an .ncu-rep:  python tools/sass_hist.py > profiles/sass_r02.md"""
import collections
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIB = os.path.join(ROOT, "icet_b200", "lib", "libicet_b200.so")
WANT = [("k_pass2", r"k_pass2ILi12ELi4"), ("k_huge_walk", r"k_huge_walk"), ("k_pass<scan2> (EXACT_PASS)", r"k_passILb1ELi12ELi4ELi2ELi1"),
        ("k_loop_cluster", r"k_loop_cluster"), ("k_loop<4>", r"k_loopILi4"), ("k_scan1_bin", r"k_scan1_bin"),
        ("k_vox2", r"k_vox2"), ("k_huge_hist", r"k_huge_hist")]
GROUPS = [("fp32 arithmetic", r"^(FADD|FMUL|FFMA|FMNMX|FSEL|FSET|FSETP|FCHK|FRND)"), ("fp64", r"^(DADD|DMUL|DFMA|DSETP|DMNMX)"),
          ("special function (MUFU)", r"^MUFU"), ("conversions", r"^(F2I|I2F|F2F|I2I|FRND|F2FP|I2FP)"),
          ("integer / logic", r"^(IADD|IADD3|IMAD|LOP3|LOP|SHF|SHL|SHR|LEA|ISETP|IMNMX|VIMNMX|SEL|POPC|FLO|BREV|PRMT|IABS|VABSDIFF|BMSK|SGXT|PLOP3|P2R|R2P)"),
          ("moves", r"^(MOV|S2R|CS2R|S2UR|R2UR|UMOV)"), ("global / local loads", r"^(LDG|LD\b|LDL|LDC|ULDC)"),
          ("global / local stores", r"^(STG|ST\b|STL)"), ("shared memory", r"^(LDS|STS|LDSM|ATOMS)"),
          ("atomics / reductions to L2", r"^(RED|ATOMG|ATOM\b)"), ("warp collectives", r"^(SHFL|VOTE|MATCH|REDUX)"),
          ("barriers / fences / cluster / mbarrier", r"^(BAR|MEMBAR|ERRBAR|CCTL|UCGABAR|WARPSYNC|NANOSLEEP|ACQBULK|DEPBAR|SYNCS|FENCE)"),
          ("control flow", r"^(BRA|BSSY|BSYNC|EXIT|CALL|RET|BRX|JMP|BREAK|YIELD|NOP|BPT)"),
          ("tensor core / TMA (UTC*MMA, UTMA*, HMMA)", r"^(UTC|UTMA|UBLKCP|HMMA|LDTM|STTM)")]


def main():
    sass = subprocess.run(["cuobjdump", "-sass", LIB], capture_output=True, text=True).stdout
    funcs, cur = {}, None
    for ln in sass.splitlines():
        m = re.search(r"Function : (\S+)", ln)
        if m:
            cur = m.group(1)
            funcs[cur] = []
            continue
        m = re.match(r"\s+/\*[0-9a-f]+\*/\s+(@!?U?P\d+\s+)?([A-Z0-9_.]+)", ln)
        if m and cur:
            funcs[cur].append(m.group(2))
    print("# SASS instruction histograms (static; `cuobjdump -sass icet_b200/lib/libicet_b200.so`, tools/sass_hist.py)\n")
    print("Static counts of the compiled code (all paths, unrolled copies included), NOT executed instructions -- those are in "
          "the ncu summaries (`smsp__inst_executed.sum`).  No `UTC*MMA`: the path has no dense contraction (SURVEY.md 8d).  "
          "`UBLKCP.S.G` = the bulk asynchronous copies (TMA engine, `cp.async.bulk`) that stage the scan-2 tiles, `SYNCS.*` = "
          "their mbarrier arrive / try_wait; `UCGABAR_*` = thread-block-cluster barriers of `k_loop_cluster`.\n")
    for title, pat in WANT:
        name = next((f for f in funcs if re.search(pat, f)), None)
        if not name:
            continue
        ops = funcs[name]
        base = collections.Counter(o.split(".")[0] for o in ops)
        print("## %s  (`%s`, %d SASS instructions)\n" % (title, name[:70], len(ops)))
        print("| group | instructions | share | top mnemonics |")
        print("|---|---|---|---|")
        left = dict(base)
        for g, rx in GROUPS:
            sel = {k: v for k, v in left.items() if re.match(rx, k)}
            for k in sel:
                left.pop(k)
            n = sum(sel.values())
            if n:
                top = ", ".join("%s %d" % kv for kv in sorted(sel.items(), key=lambda kv: -kv[1])[:5])
                print("| %s | %d | %.1f %% | %s |" % (g, n, 100.0 * n / len(ops), top))
        n = sum(left.values())
        if n:
            top = ", ".join("%s %d" % kv for kv in sorted(left.items(), key=lambda kv: -kv[1])[:6])
            print("| other | %d | %.1f %% | %s |" % (n, 100.0 * n / len(ops), top))
        full = collections.Counter(ops)
        extra = {k: v for k, v in full.items() if k.startswith(("MUFU", "UCGABAR", "RED", "ATOMG", "MATCH", "UBLKCP", "SYNCS"))}
        if extra:
            print("\nDetail: " + ", ".join("`%s` %d" % kv for kv in sorted(extra.items())))
        print()


if __name__ == "__main__":
    main()
