"""Summarise the ncu captures brought back from the GPU box (gpurun_out/) into tracked files under profiles/.

  python profiles/make_summary.py <round-tag> <launches.csv> <name=raw.csv> [<name=raw.csv> ...]

launches.csv : `ncu --metrics gpu__time_duration.sum --clock-control none --csv` of the bench command
raw.csv      : `ncu -i <rep> --page raw --csv` of a `--set full` capture
"""
import collections
import csv
import os
import shutil
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
KEYS = [
    ("gpu__time_duration.sum", "duration"),
    ("launch__grid_size", "grid"), ("launch__block_size", "block"), ("launch__registers_per_thread", "regs/thread"),
    ("launch__occupancy_limit_shared_mem", "occupancy limit (smem), CTAs"),
    ("launch__occupancy_limit_registers", "occupancy limit (regs), CTAs"),
    ("sm__warps_active.avg.pct_of_peak_sustained_active", "warps active % of peak"),
    ("smsp__issue_active.avg.pct_of_peak_sustained_active", "issue slots busy %"),
    ("sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active", "FP64 pipe active %"),
    ("sm__inst_executed_pipe_fp64.sum", "FP64 warp instructions"),
    ("smsp__inst_executed.sum", "warp instructions"),
    ("l1tex__data_pipe_lsu_wavefronts_mem_shared.sum", "shared-memory wavefronts"),
    ("l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "shared-memory bank conflicts"),
    ("dram__bytes_read.sum", "DRAM read"), ("dram__bytes_write.sum", "DRAM write"),
    ("gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "DRAM throughput % of peak"),
    ("lts__t_sector_hit_rate.pct", "L2 hit rate %"),
    ("smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio", "stall short_scoreboard (warps/issue)"),
    ("smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio", "stall long_scoreboard"),
    ("smsp__average_warps_issue_stalled_wait_per_issue_active.ratio", "stall wait (fixed latency)"),
    ("smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio", "stall barrier"),
    ("smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio", "stall math pipe throttle"),
]


def launches(path, out):
    hdr, rows = None, []
    for r in csv.reader(open(path, errors="replace")):
        if r and r[0] == "ID":
            hdr = r
        elif hdr and r and r[0].isdigit():
            rows.append(r)
    ix = {h: i for i, h in enumerate(hdr)}
    agg = collections.OrderedDict()
    for r in rows:
        k = r[ix["Kernel Name"]].split("(")[0].replace("void ", "")
        agg.setdefault(k, []).append(float(r[ix["Metric Value"]].replace(",", "")))
    mine = {k: v for k, v in agg.items() if "aceb200::" in k}
    tot = sum(sum(v) for v in mine.values())
    out.write("| kernel | launches | mean duration (us) | share of the path's kernel time |\n|---|---|---|---|\n")
    for k, v in agg.items():
        share = f"{100 * sum(v) / tot:.1f} %" if k in mine else "(not part of the path)"
        out.write(f"| `{k}` | {len(v)} | {sum(v) / len(v) / 1e3:.1f} | {share} |\n")


def raw(path, out):
    rows = list(csv.reader(open(path, errors="replace")))
    hdr, units = rows[0], rows[1]
    ix = {h: i for i, h in enumerate(hdr)}
    for r in rows[2:]:
        out.write(f"\n**`{r[ix['Kernel Name']]}`**\n\n| metric | value |\n|---|---|\n")
        for key, label in KEYS:
            if key in ix:
                out.write(f"| {label} (`{key}`) | {r[ix[key]]} {units[ix[key]]} |\n")


def main():
    tag, lpath = sys.argv[1], sys.argv[2]
    shutil.copy(lpath, os.path.join(HERE, f"{tag}_launches.csv"))
    with open(os.path.join(HERE, f"{tag}_summary.md"), "w") as out:
        out.write(f"# ncu summary, {tag}\n\nCommand: `python bench.py --steps 2 --warmup 3 --no-cpu` under "
                  "`ncu --metrics gpu__time_duration.sum --clock-control none` (launch list; cold-cache, serialised: "
                  "compare shares, not absolutes) and `ncu --set full --clock-control none --import-source on` on "
                  "`bench.py --envs 200000 --steps 1 --warmup 3` (per-kernel sections).\n\n## Launch list\n\n")
        launches(lpath, out)
        out.write("\n## `--set full` captures (2e5 environments x 40 neighbours per launch)\n")
        for spec in sys.argv[3:]:
            name, path = spec.split("=", 1)
            out.write(f"\n### {name}\n")
            raw(path, out)
            shutil.copy(path, os.path.join(HERE, f"{tag}_{name}_raw.csv"))


if __name__ == "__main__":
    main()
