"""Turn the ncu outputs of a bench run (gpurun_out/) into the tracked summaries under profiles/.
usage: python tools/summarise_profiles.py <launch_csv> <tag> [<name>=<ncu-rep | raw-page csv> ...]"""
import csv
import json
import subprocess
import sys
from collections import defaultdict


def launch_shares(path, out_md, cmd):
    rows = []
    with open(path) as fh:
        lines = [ln for ln in fh if not ln.startswith("==")]
    rd = csv.DictReader(lines)
    for r in rd:
        if r.get("Metric Name") == "gpu__time_duration.sum":
            v = float(r["Metric Value"].replace(",", ""))
            unit = r.get("Metric Unit", "ns")
            v = v * {"ns": 1e-3, "us": 1.0, "usecond": 1.0, "nsecond": 1e-3, "ms": 1e3, "msecond": 1e3}.get(unit, 1e-3)
            rows.append((r["Kernel Name"], v))
    agg = defaultdict(lambda: [0, 0.0])
    for k, v in rows:
        k = k.split("(")[0]
        agg[k][0] += 1
        agg[k][1] += v
    tot = sum(v[1] for v in agg.values())
    with open(out_md, "w") as fh:
        fh.write("# Kernel shares of the bench step (ncu launch list)\n\nCommand (on the B200 box): `%s`\n(times are cold-cache and serialised, compare SHARES).\n\n" % cmd)
        fh.write("| kernel | launches | total ms | avg us | share |\n|---|---:|---:|---:|---:|\n")
        for k, (c, t) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
            fh.write("| `%s` | %d | %.3f | %.1f | %.1f%% |\n" % (k, c, t / 1e3, t / c, 100 * t / tot))


WANT = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
        "lts__throughput.avg.pct_of_peak_sustained_elapsed", "lts__t_sector_hit_rate.pct", "l1tex__t_sector_hit_rate.pct",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "launch__registers_per_thread", "launch__grid_size", "launch__block_size",
        "smsp__issue_active.avg.pct_of_peak_sustained_active", "smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio", "smsp__average_warps_issue_stalled_membar_per_issue_active.ratio",
        "l1tex__data_pipe_lsu_wavefronts.avg.pct_of_peak_sustained_elapsed", "l1tex__m_xbar2l1tex_read_bytes.sum", "l1tex__m_xbar2l1tex_read_bytes.sum.per_second"]


def ncu_summary(rep):
    # a raw page already exported on the GPU box (the .ncu-rep files exceed what gpurun brings back) or a report file
    out = open(rep).read() if rep.endswith(".csv") else subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(out.splitlines()))
    hdr, units = rows[0], rows[1]
    res = []
    for vals in rows[2:]:
        d = {"kernel": vals[hdr.index("Kernel Name")][:70]}
        for w in WANT:
            if w in hdr:
                d[w] = "%s %s" % (vals[hdr.index(w)], units[hdr.index(w)])
        res.append(d)
    return res


if __name__ == "__main__":
    csv_path, tag = sys.argv[1], sys.argv[2]
    launch_shares(csv_path, "profiles/%s_launch_shares.md" % tag,
                  "ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file <csv> python bench.py --steps 2 --warmup 3 --no-cpu")
    summ = {}
    for a in sys.argv[3:]:
        name, rep = a.split("=")
        summ[name] = ncu_summary(rep)
    if summ:
        json.dump(summ, open("profiles/%s_ncu_full_summary.json" % tag, "w"), indent=1)
