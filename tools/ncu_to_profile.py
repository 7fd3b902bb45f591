#!/usr/bin/env python3
"""Turn gpurun_out/*.ncu-rep / launches_*.csv into the tracked summaries under profiles/.

  tools/ncu_to_profile.py <tag> [--rep gpurun_out/prof_<tag>.ncu-rep] [--launches gpurun_out/launches_<tag>.csv]

Writes profiles/<tag>_ncu_summary.md (+ .json) and refreshes profiles/latest_ncu.json, which bench.py reads for
roofline.traffic (DRAM bytes per launch of the dominant kernel, from the `ncu --set full` capture)."""
import csv
import io
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
KEYS = {
    "gpu__time_duration.sum": "time", "dram__bytes_read.sum": "dram_read", "dram__bytes_write.sum": "dram_write",
    "launch__registers_per_thread": "regs", "launch__grid_size": "grid", "launch__block_size": "block",
    "launch__occupancy_limit_registers": "occ_lim_regs", "launch__occupancy_limit_shared_mem": "occ_lim_smem",
    "sm__warps_active.avg.pct_of_peak_sustained_active": "warps_active_pct",
    "sm__throughput.avg.pct_of_peak_sustained_elapsed": "sm_throughput_pct",
    "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed": "dram_throughput_pct",
    "l1tex__throughput.avg.pct_of_peak_sustained_active": "l1tex_throughput_pct",
    "lts__throughput.avg.pct_of_peak_sustained_elapsed": "l2_throughput_pct",
    "smsp__inst_executed.sum": "warp_inst", "smsp__issue_active.avg.pct_of_peak_sustained_active": "issue_active_pct",
    "smsp__thread_inst_executed_per_inst_executed.ratio": "threads_per_inst",
    "l1tex__t_sector_hit_rate.pct": "l1_hit_pct", "lts__t_sector_hit_rate.pct": "l2_hit_pct",
    "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum": "smem_bank_conflicts",
    "sm__pipe_alu_cycles_active.avg.pct_of_peak_sustained_active": "alu_pct",
    "sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active": "fma_pct",
    "sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active": "xu_pct",
    "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active": "lsu_pct",
}
UNIT_SCALE = {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "ns": 1e-3, "us": 1.0, "ms": 1e3, "s": 1e6}


def read_rep(rep):
    out = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], stdout=subprocess.PIPE, text=True).stdout
    rows = list(csv.reader(io.StringIO(out)))
    hdr, units = rows[0], rows[1]
    idx = {h: i for i, h in enumerate(hdr)}
    kernels = {}
    for r in rows[2:]:
        name = r[idx["Kernel Name"]].split("(")[0].replace("void ", "").split("<")[0].split("::")[-1]
        rec = {}
        for k, short in KEYS.items():
            if k in idx:
                try:
                    v = float(r[idx[k]].replace(",", ""))
                except ValueError:
                    continue
                rec[short] = v * UNIT_SCALE.get(units[idx[k]], 1)
        kernels.setdefault(name, []).append(rec)
    return kernels


def main():
    tag = sys.argv[1]
    rep = sys.argv[sys.argv.index("--rep") + 1] if "--rep" in sys.argv else os.path.join(ROOT, "gpurun_out", "prof_%s.ncu-rep" % tag)
    launches = sys.argv[sys.argv.index("--launches") + 1] if "--launches" in sys.argv else os.path.join(ROOT, "gpurun_out", "launches_%s.csv" % tag)
    os.makedirs(os.path.join(ROOT, "profiles"), exist_ok=True)
    summary = {"tag": tag, "source": os.path.basename(rep), "kernels": {}}
    md = ["# ncu summary `%s`" % tag, "",
          "`ncu --set full --clock-control none --import-source on` on `python bench.py --steps 2 --warmup 3 --frames 128 --hot-only`"
          " (1 B200, one wave of 128 frames per launch). Times are per launch in microseconds, bytes per launch.", ""]
    if os.path.exists(rep):
        ks = read_rep(rep)
        md.append("| kernel | captures | time us | DRAM read | DRAM write | regs | warps active % | SM thr % | L1/TEX % | L2 % | DRAM % | issue active % | ALU pipe % | FMA pipe % | warp inst | L1 hit % | L2 hit % | smem conflicts |")
        md.append("|---|---|---|---|---|---|---|---|---|---|---|---|---|---|---|---|---|---|")
        for name, recs in ks.items():
            avg = {k: sum(r.get(k, 0) for r in recs) / len(recs) for k in recs[0]}
            avg["captures"] = len(recs)
            avg["dram_bytes_per_launch"] = avg.get("dram_read", 0) + avg.get("dram_write", 0)
            summary["kernels"][name] = avg
            md.append("| %s | %d | %.1f | %.2f MB | %.2f MB | %d | %.1f | %.1f | %.1f | %.1f | %.1f | %.1f | %.1f | %.1f | %.2f M | %.1f | %.1f | %.0f |" % (
                name, len(recs), avg.get("time", 0), avg.get("dram_read", 0) / 1e6, avg.get("dram_write", 0) / 1e6, avg.get("regs", 0),
                avg.get("warps_active_pct", 0), avg.get("sm_throughput_pct", 0), avg.get("l1tex_throughput_pct", 0),
                avg.get("l2_throughput_pct", 0), avg.get("dram_throughput_pct", 0), avg.get("issue_active_pct", 0),
                avg.get("alu_pct", 0), avg.get("fma_pct", 0), avg.get("warp_inst", 0) / 1e6, avg.get("l1_hit_pct", 0), avg.get("l2_hit_pct", 0), avg.get("smem_bank_conflicts", 0)))
        md.append("")
    if os.path.exists(launches):
        rows = [r for r in csv.reader(open(launches)) if len(r) > 5]
        hdr = next((r for r in rows if "Kernel Name" in r), None)
        if hdr:
            ik, iv, iu = hdr.index("Kernel Name"), hdr.index("Metric Value"), hdr.index("Metric Unit")
            tot = {}
            for r in rows:
                if r is hdr or len(r) <= iv:
                    continue
                try:
                    v = float(r[iv].replace(",", "")) * UNIT_SCALE.get(r[iu], 1)
                except ValueError:
                    continue
                name = r[ik].split("(")[0].replace("void ", "").split("<")[0].split("::")[-1]
                t = tot.setdefault(name, [0, 0.0])
                t[0] += 1
                t[1] += v
            all_t = sum(v[1] for v in tot.values())
            md += ["## launch list (`--metrics gpu__time_duration.sum`, cold-cache serialised: compare shares)", "",
                   "| kernel | launches | total us | share % | avg us |", "|---|---|---|---|---|"]
            summary["launch_shares"] = {}
            for name, (n, t) in sorted(tot.items(), key=lambda kv: -kv[1][1]):
                md.append("| %s | %d | %.1f | %.1f | %.2f |" % (name, n, t, 100 * t / all_t, t / n))
                summary["launch_shares"][name] = {"launches": n, "total_us": t, "share": t / all_t}
            md.append("")
    open(os.path.join(ROOT, "profiles", "%s_ncu_summary.md" % tag), "w").write("\n".join(md) + "\n")
    json.dump(summary, open(os.path.join(ROOT, "profiles", "%s_ncu_summary.json" % tag), "w"), indent=1)
    latest = {k: {"dram_bytes_per_launch": v["dram_bytes_per_launch"], "time_us": v.get("time"), "warp_inst": v.get("warp_inst"),
                  "issue_active_pct": v.get("issue_active_pct"), "alu_pct": v.get("alu_pct"), "fma_pct": v.get("fma_pct")}
              for k, v in summary["kernels"].items()}
    latest["_tag"] = tag
    latest["_commit"] = sys.argv[sys.argv.index("--commit") + 1] if "--commit" in sys.argv else subprocess.run(
        ["git", "-C", ROOT, "rev-parse", "--short", "HEAD"], stdout=subprocess.PIPE, text=True).stdout.strip()
    latest["_config"] = sys.argv[sys.argv.index("--config") + 1] if "--config" in sys.argv else "kitti"
    latest["_frames_per_launch"] = int(sys.argv[sys.argv.index("--frames-per-launch") + 1]) if "--frames-per-launch" in sys.argv else 128
    json.dump(latest, open(os.path.join(ROOT, "profiles", "latest_ncu.json"), "w"), indent=1)
    print("\n".join(md))


if __name__ == "__main__":
    main()
