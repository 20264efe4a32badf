"""Turn ncu exports into the small summaries committed under profiles/.

    python tools/ncu_summary.py launches  gpurun_out/launches.csv   profiles/rNN_launches.md
    python tools/ncu_summary.py full      gpurun_out/prof.ncu-rep   profiles/rNN_kernels.md [--json out.json]
"""
import collections
import csv
import json
import subprocess
import sys

KEYS = [
    ("gpu__time_duration.sum", "duration"),
    ("dram__bytes_read.sum", "dram read"),
    ("dram__bytes_write.sum", "dram write"),
    ("gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "dram % of peak"),
    ("dram__cycles_active.avg.pct_of_peak_sustained_elapsed", "dram cycles active %"),
    ("lts__t_sector_hit_rate.pct", "L2 hit %"),
    ("l1tex__t_sector_hit_rate.pct", "L1 hit %"),
    ("sm__warps_active.avg.pct_of_peak_sustained_active", "achieved occupancy %"),
    ("sm__throughput.avg.pct_of_peak_sustained_elapsed", "SM throughput %"),
    ("launch__registers_per_thread", "regs/thread"),
    ("launch__shared_mem_per_block_dynamic", "dyn smem/CTA"),
    ("launch__grid_size", "grid"),
    ("launch__block_size", "block"),
    ("launch__waves_per_multiprocessor", "waves/SM"),
    ("smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio", "stall long_scoreboard"),
    ("smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio", "stall barrier"),
    ("smsp__average_warps_issue_stalled_lg_throttle_per_issue_active.ratio", "stall lg_throttle"),
    ("smsp__average_warps_issue_stalled_membar_per_issue_active.ratio", "stall membar"),
    ("smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio", "stall short_scoreboard"),
]


def launches(src, dst):
    rows = list(csv.reader(open(src)))
    hdr = [i for i, r in enumerate(rows) if r and r[0] == "ID"][0]
    h = rows[hdr]
    ki, vi, ui = h.index("Kernel Name"), h.index("Metric Value"), h.index("Metric Unit")
    agg = collections.OrderedDict()
    for r in rows[hdr + 1:]:
        if len(r) <= vi:
            continue
        v = float(r[vi].replace(",", ""))
        u = r[ui]
        v = v / 1e6 if u in ("ns", "nsecond") else (v / 1e3 if u in ("us", "usecond") else v)
        agg.setdefault(r[ki], []).append(v)
    tot = sum(sum(v) for v in agg.values())
    out = ["| kernel | launches | total ms | share | avg ms |", "|---|---:|---:|---:|---:|"]
    for k, v in sorted(agg.items(), key=lambda kv: -sum(kv[1])):
        out.append(f"| `{k[:90]}` | {len(v)} | {sum(v):.3f} | {100 * sum(v) / tot:.1f}% | {sum(v) / len(v):.4f} |")
    out.append(f"\ntotal device time in the listed launches: {tot:.3f} ms "
               "(ncu serialises launches and runs them cold: compare shares, not absolutes)")
    open(dst, "w").write("\n".join(out) + "\n")
    print("\n".join(out))


def full(src, dst, json_out=None):
    txt = subprocess.run(["ncu", "-i", src, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(txt.splitlines()))
    h, units = rows[0], rows[1]
    ki = h.index("Kernel Name")
    recs = []
    out = []
    for r in rows[2:]:
        rec = {"kernel": r[ki]}
        out.append(f"### `{r[ki][:100]}`\n")
        out.append("| metric | value | unit |\n|---|---:|---|")
        for key, label in KEYS:
            if key in h:
                i = h.index(key)
                rec[key] = r[i]
                out.append(f"| {label} (`{key}`) | {r[i]} | {units[i]} |")
        out.append("")
        recs.append(rec)
    open(dst, "w").write("\n".join(out) + "\n")
    if json_out:
        json.dump(recs, open(json_out, "w"), indent=1)
    print("\n".join(out))


if __name__ == "__main__":
    mode = sys.argv[1]
    if mode == "launches":
        launches(sys.argv[2], sys.argv[3])
    else:
        j = sys.argv[sys.argv.index("--json") + 1] if "--json" in sys.argv else None
        full(sys.argv[2], sys.argv[3], j)
