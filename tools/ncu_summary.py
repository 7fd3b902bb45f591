#!/usr/bin/env python3
"""Summarise an .ncu-rep (read here, no GPU needed): per-kernel key metrics, and optionally the hottest source lines.
usage: tools/ncu_summary.py gpurun_out/prof_x.ncu-rep [--source KERNEL] [--top N]"""
import csv
import io
import subprocess
import sys

WANT = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "launch__registers_per_thread",
        "launch__grid_size", "launch__block_size", "launch__occupancy_limit_registers", "launch__occupancy_limit_shared_mem",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
        "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "l1tex__throughput.avg.pct_of_peak_sustained_active",
        "lts__throughput.avg.pct_of_peak_sustained_elapsed", "smsp__inst_executed.sum",
        "smsp__issue_active.avg.pct_of_peak_sustained_active", "smsp__thread_inst_executed_per_inst_executed.ratio",
        "l1tex__t_sectors_pipe_lsu_mem_global_op_ld.sum", "l1tex__t_requests_pipe_lsu_mem_global_op_ld.sum",
        "l1tex__t_sector_hit_rate.pct", "lts__t_sector_hit_rate.pct", "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum",
        "smsp__cycles_active.avg", "sm__inst_executed_pipe_lsu.sum", "sm__inst_executed_pipe_alu.sum",
        "sm__inst_executed_pipe_fma.sum", "sm__inst_executed_pipe_fp64.sum", "sm__inst_executed_pipe_xu.sum"]


def raw(rep):
    out = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], stdout=subprocess.PIPE, text=True).stdout
    rows = list(csv.reader(io.StringIO(out)))
    hdr, units = rows[0], rows[1]
    idx = {h: i for i, h in enumerate(hdr)}
    seen = {}
    for r in rows[2:]:
        name = r[idx["Kernel Name"]].split("(")[0]
        seen.setdefault(name, []).append(r)
    for name, rs in seen.items():
        r = rs[-1]
        print("==", name, "(%d captures, last shown)" % len(rs))
        for w in WANT:
            if w in idx:
                print("   %-62s %s %s" % (w, r[idx[w]], units[idx[w]]))


def source(rep, kernel, top):
    out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--kernel-name", "regex:" + kernel, "--launch-count", "1"],
                         stdout=subprocess.PIPE, text=True).stdout
    rows = list(csv.reader(io.StringIO(out)))
    hdr = rows[1]
    ix, isrc, ismp = hdr.index("Instructions Executed"), hdr.index("Source"), hdr.index("# Samples")
    data = [r for r in rows[2:] if len(r) > ix and r[ix].isdigit()]
    tot = sum(int(r[ix]) for r in data)
    tots = sum(int(r[ismp]) for r in data)
    print("total warp instructions", tot, "samples", tots, "sass lines", len(data))
    for r in sorted(data, key=lambda r: -int(r[ismp]))[:top]:
        print("%10s %7s  %s" % (r[ix], r[ismp], r[isrc][:110]))


if __name__ == "__main__":
    rep = sys.argv[1]
    if "--source" in sys.argv:
        k = sys.argv[sys.argv.index("--source") + 1]
        top = int(sys.argv[sys.argv.index("--top") + 1]) if "--top" in sys.argv else 40
        source(rep, k, top)
    else:
        raw(rep)


def by_line(rep, kernel, top=40):
    """Aggregate executed warp instructions and stall samples per CUDA source line (needs -lineinfo)."""
    out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--kernel-name", "regex:" + kernel, "--launch-count", "1",
                          "--print-source", "cuda,sass"], stdout=subprocess.PIPE, text=True).stdout
    rows = list(csv.reader(io.StringIO(out)))
    hdr = None
    agg = {}
    cur = None
    for r in rows:
        if r and r[0] == "Line No":
            hdr = r
            ix, ismp = hdr.index("Instructions Executed"), hdr.index("# Samples")
            continue
        if hdr is None or len(r) <= ix:
            continue
        if r[0].isdigit():            # a CUDA source line header row (aggregated over its SASS)
            cur = (int(r[0]), r[1].strip()[:100])
            try:
                agg[cur] = (agg.get(cur, (0, 0))[0] + int(r[ix]), agg.get(cur, (0, 0))[1] + int(r[ismp]))
            except ValueError:
                pass
    tot_i = sum(v[0] for v in agg.values())
    tot_s = sum(v[1] for v in agg.values())
    print("kernel %s: %d warp instructions, %d samples" % (kernel, tot_i, tot_s))
    for (ln, src), (ni, ns) in sorted(agg.items(), key=lambda kv: -kv[1][1])[:top]:
        print("%5d  inst %5.1f%%  samp %5.1f%%  %s" % (ln, 100.0 * ni / max(1, tot_i), 100.0 * ns / max(1, tot_s), src))
