#!/usr/bin/env python
"""Warp-stall samples of one kernel per SOURCE LINE, from an .ncu-rep captured with --import-source on.

    python tools/ncu_lines.py rep.ncu-rep <mangled kernel name> [top N]

ncu's CSV source page is per SASS instruction without line numbers; the line table comes from `nvdisasm -g` on the
cubin of the library the capture ran (alphazero_gym_b200/lib/libazg.so must be the same build).
"""
import collections
import csv
import os
import re
import subprocess
import sys
import tempfile

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
STALLS = ["stall_barrier", "stall_long_sb", "stall_math", "stall_mio", "stall_short_sb", "stall_wait", "stall_not_selected",
          "stall_selected", "stall_lg", "stall_dispatch", "stall_no_inst", "stall_branch_resolving", "stall_sleep", "stall_membar"]


def line_table(kernel):
    with tempfile.TemporaryDirectory() as d:
        subprocess.run(["cuobjdump", "-xelf", "all", os.path.join(ROOT, "alphazero_gym_b200", "lib", "libazg.so")], cwd=d,
                       capture_output=True)
        cubin = [f for f in os.listdir(d) if f.endswith(".cubin")][0]
        sass = subprocess.run(["nvdisasm", "-g", os.path.join(d, cubin)], capture_output=True, text=True).stdout.split("\n")
    start = [i for i, l in enumerate(sass) if l.startswith(".text." + kernel + ":")][0]
    cur, table = None, {}
    for l in sass[start + 1:]:
        if l.startswith(".text."):
            break
        m = re.match(r'\s*//## File "([^"]+)", line (\d+)', l)
        if m:
            cur = (os.path.basename(m.group(1)), int(m.group(2)))
            continue
        m = re.match(r"\s*/\*([0-9a-f]{4,})\*/\s+(.*?);", l)
        if m:
            table[int(m.group(1), 16)] = cur
    return table


def main():
    rep, kernel = sys.argv[1], sys.argv[2]
    top = int(sys.argv[3]) if len(sys.argv) > 3 else 40
    table = line_table(kernel)
    out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(out.splitlines()))
    h = rows[1]
    idx = {k: i for i, k in enumerate(h)}
    data = rows[2:]
    addr = lambda r: int(r[idx["Address"]], 16) if r[idx["Address"]].startswith("0x") else int(r[idx["Address"]])
    base = addr(data[0])
    samples, insts, stalls = collections.Counter(), collections.Counter(), collections.defaultdict(collections.Counter)
    for r in data:
        ln = table.get(addr(r) - base)
        samples[ln] += int(r[idx["# Samples"]] or 0)
        insts[ln] += int(r[idx["Instructions Executed"]] or 0)
        for s in STALLS:
            if s in idx:
                stalls[ln][s[6:]] += int(r[idx[s]] or 0)
    tot = sum(samples.values())
    print(f"samples {tot}  warp instructions {sum(insts.values())}")
    allst = collections.Counter()
    for ln in stalls:
        allst.update(stalls[ln])
    print("stall totals:", ", ".join(f"{k} {100 * v / max(1, tot):.1f}%" for k, v in allst.most_common(10)))
    src = {}
    for ln, s in samples.most_common(top):
        text = ""
        if ln:
            if ln[0] not in src:
                for dp, _, fs in os.walk(ROOT):
                    if ln[0] in fs:
                        src[ln[0]] = open(os.path.join(dp, ln[0])).read().split("\n")
                        break
                else:
                    src[ln[0]] = []
            if 0 < ln[1] <= len(src[ln[0]]):
                text = src[ln[0]][ln[1] - 1].strip()[:90]
        top_st = ", ".join(f"{k} {v}" for k, v in stalls[ln].most_common(3) if v)
        print(f"{s:6d} {100 * s / max(1, tot):5.1f}%  inst {insts[ln]:9d}  {ln[0] if ln else '?'}:{ln[1] if ln else 0:<4d} {text}   [{top_st}]")


if __name__ == "__main__":
    main()
