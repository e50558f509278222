#!/usr/bin/env python
"""Turn the ncu outputs that come back in gpurun_out/ into the small tracked summaries under profiles/.

    python tools/summarize_ncu.py launches gpurun_out/<tag>_launches.csv profiles/<name>_launches.md [steps]
    python tools/summarize_ncu.py full     gpurun_out/<tag>.ncu-rep      profiles/<name>_full.md

`launches`: per-kernel share of the captured steps (ncu serialises and cold-caches every launch, so
only the SHARES are meaningful, not the absolute times).  `full`: the handful of raw-page metrics
the roofline discussion uses (DRAM bytes, tensor-pipe %, achieved occupancy, registers).
"""
import csv
import subprocess
import sys
from collections import OrderedDict


def launches(src, dst, steps):
    rows = list(csv.DictReader(l for l in open(src) if l.startswith('"')))
    agg = OrderedDict()
    total = 0.0
    for r in rows:
        if r["Metric Name"] != "gpu__time_duration.sum":
            continue
        ns = float(r["Metric Value"].replace(",", ""))
        if r["Metric Unit"] == "us":
            ns *= 1e3
        name = r["Kernel Name"].split("(")[0].replace("void ", "")
        key = (name, r["Grid Size"], r["Block Size"])
        a = agg.setdefault(key, [0, 0.0])
        a[0] += 1
        a[1] += ns
        total += ns
    by_kernel = OrderedDict()
    for (name, g, b), (n, ns) in agg.items():
        k = by_kernel.setdefault(name, [0, 0.0])
        k[0] += n
        k[1] += ns
    with open(dst, "w") as f:
        f.write("# ncu launch list (`--metrics gpu__time_duration.sum --clock-control none`)\n\n")
        f.write("source: `%s` (%d launches, %s captured reverse-diffusion steps). Times are cold-cache and\n"
                "serialised by the profiler: read the SHARE column, not the absolute.\n\n" % (src, len(rows), steps))
        f.write("| kernel | launches | total us | share |\n|---|---:|---:|---:|\n")
        for name, (n, ns) in sorted(by_kernel.items(), key=lambda kv: -kv[1][1]):
            f.write("| `%s` | %d | %.1f | %.1f %% |\n" % (name, n, ns / 1e3, 100 * ns / total))
        f.write("\n## by launch shape\n\n| kernel | grid | block | launches | mean us | share |\n|---|---|---|---:|---:|---:|\n")
        for (name, g, b), (n, ns) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
            f.write("| `%s` | %s | %s | %d | %.1f | %.1f %% |\n" % (name, g, b, n, ns / n / 1e3, 100 * ns / total))
    print("wrote", dst)


WANT = [
    "gpu__time_duration.sum", "launch__grid_size", "launch__block_size", "launch__registers_per_thread",
    "launch__shared_mem_per_block_dynamic", "launch__occupancy_limit_shared_mem", "launch__waves_per_multiprocessor",
    "dram__bytes_read.sum", "dram__bytes_write.sum", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
    "lts__t_bytes.sum", "lts__t_sector_hit_rate.pct", "l1tex__t_bytes.sum",
    "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_tensor.sum", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
    "sm__warps_active.avg.pct_of_peak_sustained_active", "sm__cycles_active.avg", "smsp__inst_executed.sum",
    "sm__inst_executed_pipe_fma.sum", "smsp__cycles_active.avg",
]


def full(src, dst):
    out = subprocess.run(["ncu", "-i", src, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(out.splitlines()))
    hdr, units, data = rows[0], rows[1], rows[2:]
    col = {h: i for i, h in enumerate(hdr)}
    with open(dst, "w") as f:
        f.write("# ncu --set full capture (`--clock-control none --import-source on`)\n\nsource: `%s`\n\n" % src)
        f.write("| metric | unit | " + " | ".join("launch %d" % i for i in range(len(data))) + " |\n")
        f.write("|---|---|" + "---:|" * len(data) + "\n")
        f.write("| kernel | | " + " | ".join("`%s`" % d[col["Kernel Name"]].split("(")[0].replace("void ", "") for d in data) + " |\n")
        for w in WANT:
            if w in col:
                f.write("| %s | %s | %s |\n" % (w, units[col[w]], " | ".join(d[col[w]] for d in data)))
    print("wrote", dst)


if __name__ == "__main__":
    if sys.argv[1] == "launches":
        launches(sys.argv[2], sys.argv[3], sys.argv[4] if len(sys.argv) > 4 else "?")
    else:
        full(sys.argv[2], sys.argv[3])
