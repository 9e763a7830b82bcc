"""Turn gpurun_out/ ncu outputs into the tracked summaries under profiles/.

    python tools/summarize_profiles.py launches gpurun_out/launches_r1.csv profiles/r1_launches.md
    python tools/summarize_profiles.py full gpurun_out/prof_r1.ncu-rep profiles/r1_ncu_full.md
"""
import collections
import csv
import re
import subprocess
import sys


def launches(src, dst):
    lines = [l for l in open(src) if not l.startswith("==")]
    rows = []
    for row in csv.DictReader(lines):
        if row.get("Metric Name") != "gpu__time_duration.sum":
            continue
        v = float(row["Metric Value"].replace(",", ""))
        unit = row["Metric Unit"]
        v = v / 1000.0 if unit == "ns" else (v * 1000.0 if unit == "ms" else v)
        rows.append((int(row["ID"]), row["Kernel Name"], v, row["Grid Size"], row["Block Size"]))
    idx = [i for i, r in enumerate(rows) if "grid_update" in r[1]]
    start, end = (idx[-2], idx[-1]) if len(idx) >= 2 else (0, len(rows))
    step = rows[start:end]
    tot = sum(r[2] for r in step)
    agg = collections.defaultdict(lambda: [0, 0.0])
    for _, n, v, _, _ in step:
        k = re.sub(r"\(.*", "", n).replace("void ", "")[:70]
        agg[k][0] += 1
        agg[k][1] += v
    with open(dst, "w") as f:
        f.write("# ncu launch list of one navigation step (`ncu --metrics gpu__time_duration.sum --clock-control none`)\n\n")
        f.write("Source: `%s` -- launches %d..%d (one full step: grid_update -> ... -> nav_logits). Times are cold-cache and\n"
                "serialised by the profiler; compare SHARES, not absolutes.\n\n" % (src, step[0][0], step[-1][0]))
        f.write("| kernel | launches | total us | share |\n|---|---:|---:|---:|\n")
        for k, (c, v) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
            f.write("| `%s` | %d | %.1f | %.1f%% |\n" % (k, c, v, 100 * v / tot))
        f.write("| **total** | %d | %.1f | 100%% |\n\n## every launch\n\n| id | kernel | grid | block | us |\n|---:|---|---|---|---:|\n" % (len(step), tot))
        for i, n, v, g, b in step:
            f.write("| %d | `%s` | %s | %s | %.1f |\n" % (i, re.sub(r"\(.*", "", n).replace("void ", "")[:60], g, b, v))


WANT = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
        "lts__throughput.avg.pct_of_peak_sustained_elapsed", "lts__t_sector_hit_rate.pct", "l1tex__throughput.avg.pct_of_peak_sustained_elapsed",
        "sm__throughput.avg.pct_of_peak_sustained_elapsed", "sm__warps_active.avg.pct_of_peak_sustained_active",
        "launch__registers_per_thread", "launch__grid_size", "launch__block_size", "sm__cycles_elapsed.avg"]


def full(src, dst):
    raw = subprocess.run(["ncu", "-i", src, "--page", "raw", "--csv"], stdout=subprocess.PIPE, text=True).stdout
    rows = list(csv.reader(raw.splitlines()))
    hdr, units, data = rows[0], rows[1], rows[2:]
    ix = {h: i for i, h in enumerate(hdr)}
    with open(dst, "w") as f:
        f.write("# ncu --set full summaries (`%s`)\n\nOne section per captured launch; values straight from `ncu --page raw`.\n" % src)
        for r in data:
            f.write("\n## %s  grid %s\n\n| metric | value | unit |\n|---|---:|---|\n" % (r[ix["Kernel Name"]][:90], r[ix["Grid Size"]]))
            for w in WANT:
                if w in ix:
                    f.write("| %s | %s | %s |\n" % (w, r[ix[w]], units[ix[w]]))


if __name__ == "__main__":
    {"launches": launches, "full": full}[sys.argv[1]](sys.argv[2], sys.argv[3])
