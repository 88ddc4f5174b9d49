#!/usr/bin/env python
"""Turn ncu outputs in gpurun_out/ into small tracked summaries under profiles/.

  python tools/summarize_ncu.py launches gpurun_out/launches_X.csv profiles/NAME.md   # per-kernel share of a step
  python tools/summarize_ncu.py full gpurun_out/X.ncu-rep profiles/NAME.md            # key metrics of a --set full capture
"""
import collections
import csv
import re
import subprocess
import sys

KEYS = ["gpu__time_duration.sum", "launch__grid_size", "launch__block_size", "launch__registers_per_thread",
        "dram__bytes_read.sum", "dram__bytes_write.sum", "dram__bytes_read.sum.per_second",
        "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "lts__throughput.avg.pct_of_peak_sustained_elapsed",
        "lts__t_sector_hit_rate.pct", "l1tex__throughput.avg.pct_of_peak_sustained_elapsed",
        "sm__throughput.avg.pct_of_peak_sustained_elapsed", "sm__warps_active.avg.pct_of_peak_sustained_active",
        "sm__inst_issued.avg.pct_of_peak_sustained_active", "smsp__inst_executed.sum",
        "sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active",
        "sm__pipe_fmaheavy_cycles_active.avg.pct_of_peak_sustained_elapsed",
        "sm__pipe_alu_cycles_active.avg.pct_of_peak_sustained_active",
        "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum",
        "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active"]


def short(name):
    return re.sub(r"\(.*", "", name).replace("void pirb::", "").replace("pirb::", "")


def launches(src, dst):
    rows = list(csv.DictReader(l for l in open(src) if not l.startswith("==")))
    per = collections.OrderedDict()
    total = 0.0
    for r in rows:
        if r.get("Metric Name") != "gpu__time_duration.sum":
            continue
        ns = float(r["Metric Value"].replace(",", ""))
        if r["Metric Unit"] == "us":
            ns *= 1e3
        k = short(r["Kernel Name"])
        per.setdefault(k, [0, 0.0])
        per[k][0] += 1
        per[k][1] += ns
        total += ns
    with open(dst, "w") as f:
        f.write("# ncu launch list summary (`--metrics gpu__time_duration.sum --clock-control none`)\n\n")
        f.write("source: `%s` — %d launches, %.1f us total (cold-cache, serialised: compare SHARES)\n\n" % (src, len(rows), total / 1e3))
        f.write("| kernel | launches | total us | share |\n|---|---:|---:|---:|\n")
        for k, (n, ns) in sorted(per.items(), key=lambda kv: -kv[1][1]):
            f.write("| `%s` | %d | %.1f | %.1f%% |\n" % (k, n, ns / 1e3, 100 * ns / total))
    print(open(dst).read())


def full(src, dst):
    out = subprocess.run(["ncu", "-i", src, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(out.splitlines()))
    hdr, units = rows[0], rows[1]
    with open(dst, "w") as f:
        f.write("# ncu --set full summary\n\nsource: `%s`\n" % src)
        for d in rows[2:]:
            rec = dict(zip(hdr, d))
            f.write("\n## `%s`  grid %s block %s\n\n| metric | value | unit |\n|---|---:|---|\n" % (
                short(rec.get("Kernel Name", "?")), rec.get("Grid Size", "?"), rec.get("Block Size", "?")))
            for i, h in enumerate(hdr):
                if h in KEYS:
                    f.write("| %s | %s | %s |\n" % (h, d[i], units[i]))
            stalls = []
            for i, h in enumerate(hdr):
                if "pcsamp_warps_issue_stalled" in h and "not_issued" not in h:
                    try:
                        stalls.append((float(d[i]), h.split("stalled_")[1]))
                    except ValueError:
                        pass
            stalls.sort(reverse=True)
            f.write("\ntop stall reasons (pc samples): " + ", ".join("%s %d" % (n, v) for v, n in stalls[:6]) + "\n")
    print(open(dst).read()[:3000])


if __name__ == "__main__":
    {"launches": launches, "full": full}[sys.argv[1]](sys.argv[2], sys.argv[3])
