#!/usr/bin/env python3
"""Summarise ncu output for profiles/:  launches CSV (gpu__time_duration.sum per launch) -> per-kernel share table;
`--raw file.csv` (ncu -i rep --page raw --csv) -> selected metrics per kernel."""
import csv, sys, re, collections


def launches(path):
    rows = [r for r in csv.reader(open(path, errors="replace")) if len(r) > 5]
    hdr = None
    for i, r in enumerate(rows):
        if "Kernel Name" in r:
            hdr = r; rows = rows[i + 1:]; break
    kn, mv, mu = hdr.index("Kernel Name"), hdr.index("Metric Value"), hdr.index("Metric Unit")
    agg = collections.OrderedDict()
    for r in rows:
        try:
            v = float(r[mv].replace(",", ""))
        except ValueError:
            continue
        unit = r[mu]
        ms = v * {"ns": 1e-6, "us": 1e-3, "usecond": 1e-3, "nsecond": 1e-6, "ms": 1.0, "msecond": 1.0, "s": 1e3, "second": 1e3}.get(unit, 1e-6)
        name = re.sub(r"\(.*", "", r[kn])
        a = agg.setdefault(name, [0, 0.0]); a[0] += 1; a[1] += ms
    tot = sum(a[1] for a in agg.values())
    print("| kernel | launches | total ms | share | avg ms |\n|---|---|---|---|---|")
    for name, (n, ms) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
        print("| %s | %d | %.3f | %.1f%% | %.4f |" % (name, n, ms, 100 * ms / tot, ms / n))
    print("\ntotal kernel time %.1f ms over %d launches (cold-cache, serialised by ncu: compare shares, not absolutes)" % (tot, sum(a[0] for a in agg.values())))


def raw(path, pats):
    rows = list(csv.reader(open(path, errors="replace")))
    hdr = rows[0]; units = rows[1]
    kn = hdr.index("Kernel Name")
    cols = [i for i, h in enumerate(hdr) if any(re.search(p, h) for p in pats)]
    for r in rows[2:]:
        if len(r) <= kn:
            continue
        print("## " + r[kn][:120])
        for i in cols:
            print("  %-70s %s %s" % (hdr[i], r[i], units[i]))


if __name__ == "__main__":
    if sys.argv[1] == "--raw":
        pats = sys.argv[3:] or [r"gpu__time_duration.sum", r"dram__bytes_(read|write)\.sum$", r"sm__throughput.avg.pct", r"sm__inst_executed_pipe_fp64", r"sm__pipe_fp64_cycles_active", r"smsp__inst_executed_pipe_fp64",
                                r"sm__warps_active.avg.pct_of_peak", r"launch__registers_per_thread", r"launch__occupancy_limit", r"lts__t_bytes.sum$", r"l1tex__t_bytes.sum$",
                                r"sm__pipe_tensor.*cycles_active", r"smsp__issue_active.avg.pct", r"sm__inst_executed_pipe_tensor", r"gpu__dram_throughput", r"lts__t_sectors_op_red", r"smsp__average_warp.*stall", r"achieved_occupancy", r"sm__cycles_active.avg$", r"dmma", r"lts__throughput"]
        raw(sys.argv[2], pats)
    else:
        launches(sys.argv[1])
