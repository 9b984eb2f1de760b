#!/usr/bin/env python
"""Turn ncu outputs into the small text summaries committed under profiles/.

    python tools/ncu_summary.py launches <launches.csv> <out.md>         # gpu__time_duration per launch -> per-kernel table
    python tools/ncu_summary.py full <report.ncu-rep> <out.md> [traffic.json]   # --set full capture -> key metrics (+ traffic.json)
"""
import collections
import csv
import json
import subprocess
import sys

KEYS = [
    "gpu__time_duration.sum", "launch__grid_size", "launch__block_size", "launch__registers_per_thread",
    "launch__occupancy_limit_registers", "launch__occupancy_limit_shared_mem", "launch__occupancy_limit_warps",
    "sm__warps_active.avg.pct_of_peak_sustained_active", "smsp__issue_active.avg.pct_of_peak_sustained_active",
    "smsp__inst_executed.sum", "smsp__thread_inst_executed_per_inst_executed.ratio",
    "dram__bytes_read.sum", "dram__bytes_write.sum", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
    "lts__t_sector_hit_rate.pct", "l1tex__t_sector_hit_rate.pct",
    "sm__throughput.avg.pct_of_peak_sustained_elapsed",
    "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum",
    "TPC.TriageCompute.sm__pipe_tensor_cycles_active_realtime.avg.pct_of_peak_sustained_elapsed",
    "sm__inst_executed_pipe_tc.sum.pct_of_peak_sustained_active",
    "l1tex__data_pipe_tc_wavefronts_mem_shared.sum.pct_of_peak_sustained_elapsed",
    "smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_wait_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_lg_throttle_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_mio_throttle_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio",
]


def launches(src, dst):
    rows = [r for r in csv.reader(open(src)) if len(r) > 5]
    hdr, agg, order = None, collections.defaultdict(list), []
    for r in rows:
        if r[0] == "ID":
            hdr = r
            continue
        if hdr is None:
            continue
        d = dict(zip(hdr, r))
        name = d["Kernel Name"].split("(")[0].replace("void ", "")
        if name not in agg:
            order.append(name)
        agg[name].append((float(d["Metric Value"].replace(",", "")), d["Grid Size"]))
    unit = "ns"
    tot = sum(v for vs in agg.values() for v, _ in vs)
    with open(dst, "w") as f:
        f.write(f"per-kernel device time from `ncu --metrics gpu__time_duration.sum --clock-control none` ({src})\n")
        f.write("(cold-cache, serialised launches: compare SHARES, not absolutes)\n\n")
        f.write("| kernel | launches | mean us | share of captured time | grid of first launch |\n|---|---|---|---|---|\n")
        for name in sorted(order, key=lambda n: -sum(v for v, _ in agg[n])):
            vs = agg[name]
            s = sum(v for v, _ in vs)
            f.write(f"| {name} | {len(vs)} | {s / len(vs) / 1e3:.1f} | {s / tot:.3f} | {vs[0][1]} |\n")
        f.write(f"\ntotal captured: {tot / 1e6:.3f} ms over {sum(len(v) for v in agg.values())} launches ({unit})\n")


def full(rep, dst, traffic_json=None):
    raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(raw.splitlines()))
    hdr, units = rows[0], rows[1]
    traffic = {}
    with open(dst, "w") as f:
        f.write(f"key metrics of `ncu --set full --clock-control none --import-source on` capture {rep}\n\n")
        for r in rows[2:]:
            d = dict(zip(hdr, r))
            u = dict(zip(hdr, units))
            name = d["Kernel Name"].split("(")[0].replace("void ", "")
            f.write(f"## {name}   grid {d.get('Grid Size', '')} block {d.get('Block Size', '')}\n\n| metric | value | unit |\n|---|---|---|\n")
            for k in KEYS:
                if k in d:
                    f.write(f"| {k} | {d[k]} | {u[k]} |\n")
            try:
                def tobytes(key):
                    v, un = float(d[key].replace(",", "")), u[key].lower()
                    return v * {"byte": 1, "kbyte": 1e3, "mbyte": 1e6, "gbyte": 1e9}[un]
                tb = tobytes("dram__bytes_read.sum") + tobytes("dram__bytes_write.sum")
                f.write(f"| dram read+write per launch | {tb / 1e6:.2f} | MB |\n")
                traffic[name.split("::")[-1].split("<")[0]] = tb
            except Exception:
                pass
            f.write("\n")
        # hottest source lines (stall samples) of the first kernel
        src = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
        srows = list(csv.reader(src.splitlines()))
        if len(srows) > 3:
            body = srows[2:]
            try:
                tot_s = sum(int(r[4]) for r in body)
                tot_i = sum(int(r[5]) for r in body)
                f.write(f"### SASS hot spots of {srows[0][1][:80]} (samples {tot_s}, warp instructions {tot_i})\n\n")
                f.write("| # | SASS | warp instr | avg active threads | stall samples |\n|---|---|---|---|---|\n")
                for i, r in sorted(enumerate(body), key=lambda ir: -int(ir[1][4]))[:14]:
                    f.write(f"| {i} | `{r[1].strip()[:70]}` | {r[5]} | {r[8]} | {r[4]} |\n")
            except Exception as e:  # pragma: no cover
                f.write(f"(source page not parsed: {e})\n")
    if traffic_json:
        try:
            old = json.load(open(traffic_json))
        except Exception:
            old = {}
        old.update(traffic)
        json.dump(old, open(traffic_json, "w"), indent=1, sort_keys=True)


if __name__ == "__main__":
    if sys.argv[1] == "launches":
        launches(sys.argv[2], sys.argv[3])
    else:
        full(sys.argv[2], sys.argv[3], sys.argv[4] if len(sys.argv) > 4 else None)
