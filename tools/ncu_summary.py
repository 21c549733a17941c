"""Summarise an .ncu-rep (read here, no GPU needed) into the short table committed under profiles/.

    python tools/ncu_summary.py gpurun_out/prof.ncu-rep [--out profiles/name.md] [--title "..."]
"""
import argparse
import csv
import io
import subprocess

KEYS = [
    ("gpu__time_duration.sum", "duration"),
    ("launch__grid_size", "grid"),
    ("launch__block_size", "block"),
    ("launch__registers_per_thread", "registers/thread"),
    ("launch__shared_mem_per_block_dynamic", "dynamic smem/block"),
    ("sm__warps_active.avg.pct_of_peak_sustained_active", "achieved occupancy %"),
    ("dram__bytes_read.sum", "DRAM read"),
    ("dram__bytes_write.sum", "DRAM written"),
    ("dram__bytes_read.sum.per_second", "DRAM read rate"),
    ("gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "DRAM throughput % of peak"),
    ("lts__t_sector_hit_rate.pct", "L2 hit rate %"),
    ("l1tex__t_sector_hit_rate.pct", "L1 hit rate %"),
    ("sm__throughput.avg.pct_of_peak_sustained_elapsed", "SM throughput %"),
    ("sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active", "fp64 pipe active % (DFMA)"),
    ("sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active", "XU pipe % (F2F.F64.F32; on this chip the counter also moves with DMMA)"),
    ("sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active", "tensor pipe active % (all sub-pipes)"),
    ("sm__pipe_tensor_subpipe_dmma_cycles_active.avg.pct_of_peak_sustained_active", "fp64 tensor sub-pipe (DMMA) active %"),
    ("sm__pipe_tensor_subpipe_imma_cycles_active.avg.pct_of_peak_sustained_active", "int8 tensor sub-pipe (IMMA) active %"),
    ("smsp__issue_active.avg.pct_of_peak_sustained_active", "issue slots busy %"),
    ("smsp__inst_executed.sum", "warp instructions"),
    ("smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio", "stall long_scoreboard (per issue)"),
    ("smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio", "stall short_scoreboard"),
    ("smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio", "stall math_pipe_throttle"),
    ("smsp__average_warps_issue_stalled_mio_throttle_per_issue_active.ratio", "stall mio_throttle"),
    ("smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio", "stall barrier"),
    ("smsp__average_warps_issue_stalled_wait_per_issue_active.ratio", "stall wait"),
    ("smsp__average_warps_issue_stalled_not_selected_per_issue_active.ratio", "stall not_selected"),
]


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("rep")
    ap.add_argument("--out")
    ap.add_argument("--title", default="")
    ap.add_argument("--note", default="")
    args = ap.parse_args()
    raw = subprocess.run(["ncu", "-i", args.rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(raw)))
    hdr, units, data = rows[0], rows[1], rows[2:]
    ki = hdr.index("Kernel Name")
    lines = [f"# {args.title or args.rep}", "",
             f"Source: `{args.rep}` (`ncu --set full --clock-control none --import-source on`), read with "
             "`ncu -i ... --page raw --csv`. Per-launch values; ncu replays each launch ~40x with cold caches, so "
             "durations here are NOT bench numbers.", ""]
    if args.note:
        lines += [args.note, ""]
    names = [r[ki].split("(")[0].replace("void ", "") for r in data]
    lines.append("| metric | " + " | ".join(f"launch {i}: `{n}`" for i, n in enumerate(names)) + " |")
    lines.append("|---|" + "---|" * len(data))
    for key, label in KEYS:
        if key not in hdr:
            continue
        i = hdr.index(key)
        vals = []
        for r in data:
            v = r[i]
            try:
                f = float(v.replace(",", ""))
                v = f"{f:,.3f}".rstrip("0").rstrip(".") if abs(f) < 1e6 else f"{f:,.0f}"
            except ValueError:
                pass
            vals.append(f"{v} {units[i]}".strip())
        lines.append(f"| {label} (`{key}`) | " + " | ".join(vals) + " |")
    text = "\n".join(lines) + "\n"
    if args.out:
        open(args.out, "w").write(text)
    print(text)


if __name__ == "__main__":
    main()
