"""Turn gpurun_out/ ncu outputs into the tracked summaries under profiles/.

    python tools/summarize_profiles.py launches gpurun_out/launches_r1.csv profiles/r1_launches.md
    python tools/summarize_profiles.py full gpurun_out/prof_r1.ncu-rep profiles/r1_ncu_full.md
    python tools/summarize_profiles.py step gpurun_out/prof_r1c_all_raw.csv profiles/r1_ncu_full_v4_step.md profiles/r1_traffic.json gpurun_out/prof_r1c_pool.ncu-rep
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


def _bytes(v, u):
    return float(v.replace(",", "")) * {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}.get(u, 1)


def step(src_csv, dst, traffic_json=None, pool_rep=None):
    """`ncu --page raw --csv` export of every launch of one step (tools/gpu_evidence.sh) -> per-launch tables + per-kernel
    totals; optionally profiles/r1_traffic.json (DRAM bytes of the GEMM launches and of the pooling kernel)."""
    import json
    rows = list(csv.reader(open(src_csv)))
    hdr, units, data = rows[0], rows[1], rows[2:]
    ix = {h: i for i, h in enumerate(hdr)}
    agg = collections.defaultdict(lambda: [0, 0.0, 0.0, 0.0])
    gem = 0.0
    with open(dst, "w") as f:
        f.write("# ncu --set full, every launch of this library in one navigation step (`ncu --page raw --csv` exported on the GPU box)\n")
        for r in data:
            name = r[ix["Kernel Name"]]
            f.write("\n## %s  grid %s\n\n| metric | value | unit |\n|---|---:|---|\n" % (name[:100], r[ix["Grid Size"]]))
            for w in WANT:
                if w in ix:
                    f.write("| %s | %s | %s |\n" % (w, r[ix[w]], units[ix[w]]))
            b = _bytes(r[ix["dram__bytes_read.sum"]], units[ix["dram__bytes_read.sum"]]) + \
                _bytes(r[ix["dram__bytes_write.sum"]], units[ix["dram__bytes_write.sum"]])
            key = re.sub(r"[<(].*", "", name).replace("void ", "").replace("gmm::", "")
            t = float(r[ix["gpu__time_duration.sum"]].replace(",", ""))
            tp = float(r[ix["sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active"]].replace(",", ""))
            agg[key][0] += 1; agg[key][1] += b; agg[key][2] += t; agg[key][3] += tp * t
            if "gemm" in name:
                gem += b
        f.write("\n## totals per kernel (one step)\n\n| kernel | launches | DRAM bytes (MB) | ncu time (us) | tensor pipe active, time-weighted (%) |\n|---|---:|---:|---:|---:|\n")
        for k, (n, b, t, tp) in sorted(agg.items(), key=lambda kv: -kv[1][2]):
            f.write("| `%s` | %d | %.1f | %.1f | %.1f |\n" % (k, n, b / 1e6, t, tp / t if t else 0))
    if traffic_json and pool_rep:
        raw = subprocess.run(["ncu", "-i", pool_rep, "--page", "raw", "--csv"], stdout=subprocess.PIPE, text=True).stdout
        pr = list(csv.reader(raw.splitlines()))
        pix = {h: i for i, h in enumerate(pr[0])}
        pb = sum(_bytes(pr[2][pix[m]], pr[1][pix[m]]) for m in ("dram__bytes_read.sum", "dram__bytes_write.sum"))
        json.dump({"source": "ncu --set full --clock-control none (tools/gpu_r2_final.sh, B=32 T=8 step): dram__bytes_read.sum + dram__bytes_write.sum",
                   "pool_bytes_per_launch": pb, "gemm_bytes_per_step": gem}, open(traffic_json, "w"), indent=1)


if __name__ == "__main__":
    {"launches": launches, "full": full, "step": step}[sys.argv[1]](*sys.argv[2:])
