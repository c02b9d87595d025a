"""Turns ncu outputs under gpurun_out/ into the committed summaries under profiles/.

    python tools/summarize_ncu.py --launches gpurun_out/launches_r1.csv --rep gpurun_out/prof_r1.ncu-rep --tag r01
"""
import argparse
import collections
import csv
import io
import subprocess

METRICS = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
           "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
           "sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active",
           "sm__throughput.avg.pct_of_peak_sustained_elapsed", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
           "lts__t_sector_hit_rate.pct", "launch__registers_per_thread", "launch__grid_size",
           "smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio",
           "smsp__average_warps_issue_stalled_wait_per_issue_active.ratio",
           "smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio"]


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--launches")
    ap.add_argument("--rep")
    ap.add_argument("--tag", default="r01")
    a = ap.parse_args()
    if a.launches:
        rows = [r for r in csv.reader(open(a.launches)) if len(r) > 10 and r[0].isdigit()]
        agg = collections.OrderedDict()
        for r in rows:
            name = r[4].split("(")[0].replace("void ", "").replace("synt::", "")
            e = agg.setdefault(name, [0, 0.0])
            e[0] += 1
            e[1] += float(r[-1]) / 1e3
        total = sum(v[1] for v in agg.values())
        with open(f"profiles/{a.tag}_launch_shares.md", "w") as f:
            f.write(f"# {a.tag}: per-kernel share of one sampling step (ncu --metrics gpu__time_duration.sum, "
                    f"--clock-control none; cold-cache, serialised -> compare SHARES)\n\n")
            f.write(f"source: `{a.launches}` ({len(rows)} launches, B=64, tools/ncu_target.py)\n\n")
            f.write("| kernel | launches | total us | share |\n|---|---:|---:|---:|\n")
            for k, (n, us) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
                f.write(f"| `{k}` | {n} | {us:.1f} | {100 * us / total:.1f}% |\n")
            f.write(f"| **total** | {len(rows)} | {total:.1f} | 100% |\n")
        with open(f"profiles/{a.tag}_launches.csv", "w") as f:
            f.write("id,kernel,block,grid,duration_ns\n")
            for r in rows:
                f.write(f"{r[0]},\"{r[4]}\",\"{r[7]}\",\"{r[8]}\",{r[-1]}\n")
    if a.rep:
        raw = subprocess.run(["ncu", "-i", a.rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
        rows = list(csv.reader(io.StringIO(raw)))
        hdr, units = rows[0], rows[1]
        idx = {m: hdr.index(m) for m in METRICS if m in hdr}
        with open(f"profiles/{a.tag}_ncu_full_metrics.csv", "w") as f:
            w = csv.writer(f)
            w.writerow(["kernel", "grid"] + [f"{m} [{units[i]}]" for m, i in idx.items()])
            for r in rows[2:]:
                w.writerow([r[hdr.index("Kernel Name")], r[hdr.index("Grid Size")]] + [r[i] for i in idx.values()])


if __name__ == "__main__":
    main()
