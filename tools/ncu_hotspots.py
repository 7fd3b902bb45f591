#!/usr/bin/env python3
"""Per-source-line summary of an `ncu --set full --import-source on` capture: share of executed warp instructions and of
stall samples per line of orb_kernels.cuh, for one kernel.  usage: tools/ncu_hotspots.py <rep> <kernel regex> [top N]"""
import csv
import io
import subprocess
import sys


def main():
    rep, kern = sys.argv[1], sys.argv[2]
    top = int(sys.argv[3]) if len(sys.argv) > 3 else 25
    out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "-k", "regex:" + kern, "--print-source", "cuda,sass"],
                         stdout=subprocess.PIPE, text=True).stdout
    rows = list(csv.reader(io.StringIO(out)))
    start = next(i for i, r in enumerate(rows) if r and r[0] == "Line No")
    hdr = rows[start]
    ix_s, ix_i = hdr.index("# Samples"), hdr.index("Instructions Executed")
    lines = []
    for r in rows[start + 1:]:
        if r and r[0] != "":
            try:
                lines.append((int(r[0]), r[1].strip(), int(r[ix_s]), int(r[ix_i])))
            except ValueError:
                pass
    ti, ts = sum(l[3] for l in lines) or 1, sum(l[2] for l in lines) or 1
    print("### %s: %d warp instructions, %d stall samples\n" % (kern, ti, ts))
    print("| line | instr % | samples % | source |\n|---|---|---|---|")
    for ln, src, s, n in sorted(lines, key=lambda t: -(t[2] / ts + t[3] / ti))[:top]:
        print("| %d | %.1f | %.1f | `%s` |" % (ln, 100.0 * n / ti, 100.0 * s / ts, src[:110].replace("|", "\\|")))
    print()


if __name__ == "__main__":
    main()
