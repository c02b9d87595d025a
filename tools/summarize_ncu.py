"""Turns ncu outputs under gpurun_out/ into the committed summaries under profiles/.

    python tools/summarize_ncu.py --launches gpurun_out/launches_r1c.csv --tag r01c [--skip N --take M]
    python tools/summarize_ncu.py --raw gpurun_out/r1c_conv_raw.csv --tag r01c_conv

--launches: the CSV of `ncu --metrics gpu__time_duration.sum[,dram__bytes_read.sum,dram__bytes_write.sum]
--clock-control none --csv` (one row per launch and metric).  Writes profiles/<tag>_launches.csv (one row per launch),
profiles/<tag>_launch_shares.md (per-kernel share of the window) and profiles/<tag>_traffic.json (average DRAM bytes
per launch of every kernel -- bench.py reads the conv entry for `roofline.traffic`).
--raw: the `--page raw --csv` export of an `ncu --set full` report; keeps the columns that matter.
"""
import argparse
import collections
import csv
import json

csv.field_size_limit(10 ** 9)

METRICS = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
           "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
           "sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active",
           "sm__throughput.avg.pct_of_peak_sustained_elapsed", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
           "smsp__issue_active.avg.pct_of_peak_sustained_active", "sm__warps_active.avg.pct_of_peak_sustained_active",
           "lts__t_sector_hit_rate.pct", "launch__registers_per_thread", "launch__grid_size", "launch__block_size",
           "launch__shared_mem_per_block_dynamic", "launch__occupancy_limit_shared_mem", "launch__occupancy_limit_registers",
           "sm__cycles_elapsed.avg", "smsp__inst_executed.sum",
           "smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio",
           "smsp__average_warps_issue_stalled_wait_per_issue_active.ratio",
           "smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio"]


def short(name):
    return name.split("(")[0].replace("void ", "").replace("synt::", "")


def to_bytes(value, unit):
    v = float(value.replace(",", ""))
    return v * {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}.get(unit, 1)


def launches(path, tag, skip, take, note):
    rows = [r for r in csv.reader(open(path)) if len(r) > 10 and r[0].isdigit()]
    per = collections.OrderedDict()          # id -> {kernel, block, grid, metrics}
    for r in rows:
        e = per.setdefault(int(r[0]), {"kernel": r[4], "block": r[7], "grid": r[8], "m": {}})
        e["m"][r[-3]] = (r[-1], r[-2])
    ids = sorted(per)[skip:]
    if take:
        ids = ids[:take]
    agg = collections.OrderedDict()
    with open(f"profiles/{tag}_launches.csv", "w") as f:
        f.write("id,kernel,block,grid,duration_ns,dram_read_bytes,dram_write_bytes\n")
        for i in ids:
            e = per[i]
            dur = float(e["m"]["gpu__time_duration.sum"][0].replace(",", ""))
            rd = to_bytes(*e["m"]["dram__bytes_read.sum"]) if "dram__bytes_read.sum" in e["m"] else None
            wr = to_bytes(*e["m"]["dram__bytes_write.sum"]) if "dram__bytes_write.sum" in e["m"] else None
            f.write(f"{i},\"{e['kernel']}\",\"{e['block']}\",\"{e['grid']}\",{dur:.0f},"
                    f"{'' if rd is None else f'{rd:.0f}'},{'' if wr is None else f'{wr:.0f}'}\n")
            a = agg.setdefault(short(e["kernel"]), [0, 0.0, 0.0, 0])
            a[0] += 1
            a[1] += dur / 1e3
            if rd is not None:
                a[2] += rd + wr
                a[3] += 1
    total = sum(v[1] for v in agg.values())
    with open(f"profiles/{tag}_launch_shares.md", "w") as f:
        f.write(f"# {tag}: per-kernel share of the profiled window (ncu --metrics gpu__time_duration.sum,dram__bytes_*.sum, "
                f"--clock-control none; cold-cache, serialised -> compare SHARES, not absolutes)\n\n")
        f.write(f"source: `{path}` (launches {ids[0]}..{ids[-1]}, {len(ids)} launches){'; ' + note if note else ''}\n\n")
        f.write("| kernel | launches | total us | share | avg DRAM MB / launch |\n|---|---:|---:|---:|---:|\n")
        for k, (n, us, by, nb) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
            f.write(f"| `{k}` | {n} | {us:.1f} | {100 * us / total:.1f}% | {'' if not nb else f'{by / nb / 1e6:.2f}'} |\n")
        f.write(f"| **total** | {len(ids)} | {total:.1f} | 100% | |\n")
    traffic = {k: {"launches": n, "avg_dram_bytes_per_launch": (by / nb if nb else None), "total_us": us}
               for k, (n, us, by, nb) in agg.items()}
    json.dump({"source": path, "window": [ids[0], ids[-1]], "kernels": traffic}, open(f"profiles/{tag}_traffic.json", "w"), indent=1)
    print(open(f"profiles/{tag}_launch_shares.md").read())


def raw(path, tag):
    rows = list(csv.reader(open(path)))
    hdr, units = rows[0], rows[1]
    idx = {m: hdr.index(m) for m in METRICS if m in hdr}
    with open(f"profiles/{tag}_ncu_full_metrics.csv", "w") as f:
        w = csv.writer(f)
        w.writerow(["kernel", "grid"] + [f"{m} [{units[i]}]" for m, i in idx.items()])
        for r in rows[2:]:
            w.writerow([short(r[hdr.index("Kernel Name")]), r[hdr.index("Grid Size")]] + [r[i] for i in idx.values()])
    print(open(f"profiles/{tag}_ncu_full_metrics.csv").read()[:3000])


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--launches")
    ap.add_argument("--raw")
    ap.add_argument("--tag", default="r01")
    ap.add_argument("--skip", type=int, default=0)
    ap.add_argument("--take", type=int, default=0)
    ap.add_argument("--note", default="")
    a = ap.parse_args()
    if a.launches:
        launches(a.launches, a.tag, a.skip, a.take, a.note)
    if a.raw:
        raw(a.raw, a.tag)


if __name__ == "__main__":
    main()
