"""Turn the raw ncu exports brought back in gpurun_out/ into the small, committed summaries under profiles/.

    python profiles/summarize.py launches gpurun_out/launches.csv profiles/r01_launches_summary.md
    python profiles/summarize.py kernel   gpurun_out/prof_main_raw.csv profiles/r01_k_bpr_main_ncu.md

`launches`: the `ncu --metrics gpu__time_duration.sum --clock-control none --csv` launch list (cold-cache and
serialised: compare SHARES, not absolutes).  `kernel`: `ncu -i x.ncu-rep --page raw --csv` of a `--set full` capture.
"""
import collections
import csv
import json
import sys

KEEP = [
    "gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "dram__bytes.sum.per_second",
    "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "lts__t_sector_hit_rate.pct",
    "lts__throughput.avg.pct_of_peak_sustained_elapsed", "l1tex__throughput.avg.pct_of_peak_sustained_active",
    "sm__throughput.avg.pct_of_peak_sustained_elapsed", "sm__warps_active.avg.pct_of_peak_sustained_active",
    "smsp__issue_active.avg.pct_of_peak_sustained_active", "launch__registers_per_thread", "launch__grid_size",
    "launch__block_size", "launch__occupancy_limit_registers", "launch__waves_per_multiprocessor",
    "smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_lg_throttle_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_wait_per_issue_active.ratio",
    "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
]


def launches(src, dst):
    lines = [l for l in open(src) if not l.startswith("==")]
    agg = collections.OrderedDict()
    for row in csv.DictReader(lines):
        v = float(row["Metric Value"].replace(",", ""))
        v = {"ns": v / 1e3, "us": v, "ms": v * 1e3, "s": v * 1e6}.get(row["Metric Unit"], v)
        a = agg.setdefault(row["Kernel Name"].split("(")[0][-90:], [0, 0.0])
        a[0] += 1
        a[1] += v
    tot = sum(a[1] for a in agg.values())
    with open(dst, "w") as f:
        f.write("| kernel | launches | avg us | total us | share |\n|---|---:|---:|---:|---:|\n")
        for k, (c, t) in sorted(agg.items(), key=lambda x: -x[1][1]):
            f.write(f"| `{k}` | {c} | {t / c:.1f} | {t:.0f} | {100 * t / tot:.1f} % |\n")
    print(open(dst).read())


def kernel(src, dst):
    rows = list(csv.reader(open(src)))
    hdr, units = rows[0], rows[1]
    out = []
    for r in rows[2:]:
        d = {"kernel": r[hdr.index("Kernel Name")]}
        for k in KEEP:
            if k in hdr:
                d[k] = (r[hdr.index(k)], units[hdr.index(k)])
        out.append(d)
    with open(dst, "w") as f:
        for d in out:
            f.write(f"### `{d['kernel'][:100]}`\n\n| metric | value | unit |\n|---|---:|---|\n")
            for k in KEEP:
                if k in d:
                    f.write(f"| {k} | {d[k][0]} | {d[k][1]} |\n")
            f.write("\n")
    print(open(dst).read())
    return out


if __name__ == "__main__":
    {"launches": launches, "kernel": kernel}[sys.argv[1]](sys.argv[2], sys.argv[3])
