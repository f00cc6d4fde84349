#!/usr/bin/env python
"""turns ncu CSV exports into the digests committed under profiles/

  python scripts/ncu_digest.py launches <launches.csv> <out_summary.txt> [--skip-setup]
  python scripts/ncu_digest.py full <raw.csv> <out_digest.csv> <out_traffic.json> <n>
"""
import csv, json, re, sys
from collections import OrderedDict

KEEP = ["Kernel Name", "gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "lts__t_sectors.sum",
        "lts__t_sector_hit_rate.pct", "sm__warps_active.avg.pct_of_peak_sustained_active", "launch__registers_per_thread",
        "launch__grid_size", "launch__block_size", "launch__shared_mem_per_block_dynamic", "smsp__issue_active.avg.pct_of_peak_sustained_active",
        "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "dram__cycles_active.avg.pct_of_peak_sustained_elapsed",
        "launch__occupancy_limit_registers", "launch__occupancy_limit_shared_mem", "sm__cycles_elapsed.avg.per_second",
        "smsp__average_warp_latency_issue_stalled_long_scoreboard.pct", "smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio", "smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_sleeping_per_issue_active.ratio", "smsp__average_warps_issue_stalled_membar_per_issue_active.ratio",
        "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "smsp__inst_executed.sum"]


def short(name):
    m = re.match(r"(?:void )?([A-Za-z0-9_:]+(?:<[^(]*>)?)", name)
    s = m.group(1) if m else name
    return s if len(s) < 70 else s[:67] + "..."


def launches(path, out, pattern=None):
    rows = [r for r in csv.reader(l for l in open(path) if l.startswith('"'))]
    hdr = rows[0]
    ik, iv = hdr.index("Kernel Name"), hdr.index("Metric Value")
    agg = OrderedDict()
    tot = 0.0
    n = 0
    for r in rows[1:]:
        ns = float(r[iv].replace(",", ""))
        k = short(r[ik])
        a = agg.setdefault(k, [0, 0.0])
        a[0] += 1
        a[1] += ns
        tot += ns
        n += 1
    with open(out, "w") as f:
        f.write("# source: %s\n# cold-cache serialised per-launch times under ncu: compare SHARES, not absolutes\n" % path)
        f.write("launches %d  total %.3f ms\n" % (n, tot / 1e6))
        for k, (c, ns) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
            f.write("%-72s n=%5d %10.3f ms %6.1f%%\n" % (k, c, ns / 1e6, 100 * ns / tot))
    print(open(out).read())


def full(path, out, traffic_out, n):
    rows = list(csv.reader(open(path)))
    hdr, units = rows[0], rows[1]
    cols = [hdr.index(k) for k in KEEP if k in hdr]
    with open(out, "w", newline="") as f:
        w = csv.writer(f)
        w.writerow([hdr[c] for c in cols])
        w.writerow([units[c] for c in cols])
        for r in rows[2:]:
            w.writerow([r[c] for c in cols])
    ir, iw, ik = hdr.index("dram__bytes_read.sum"), hdr.index("dram__bytes_write.sum"), hdr.index("Kernel Name")
    scale = {"Gbyte": 1e9, "Mbyte": 1e6, "Kbyte": 1e3, "byte": 1.0}
    per = {}
    for r in rows[2:]:
        by = float(r[ir]) * scale[units[ir]] + float(r[iw]) * scale[units[iw]]
        key = short(r[ik])
        per.setdefault(key, []).append(by)
    tj = {"n": int(n), "source": "ncu --set full --clock-control none (digest %s): dram__bytes_read.sum + dram__bytes_write.sum per launch, mean over the captured launches" % out,
          "kernels": {k: sum(v) / len(v) for k, v in per.items()}}
    json.dump(tj, open(traffic_out, "w"), indent=1)
    print(json.dumps(tj, indent=1))


if __name__ == "__main__":
    if sys.argv[1] == "launches":
        launches(sys.argv[2], sys.argv[3])
    else:
        full(*sys.argv[2:6])
