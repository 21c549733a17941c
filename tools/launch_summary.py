"""Summarise an `ncu --metrics gpu__time_duration.sum --csv` launch list into per-kernel shares of one bench step.

    python tools/launch_summary.py gpurun_out/launches.csv --passes 5 --keep 2 [--out profiles/name.md] [--title "..."]

`--passes`: hot-path passes the profiled command ran (warm-up + timed); `--keep`: how many of the LAST passes to average.
ncu serialises launches and runs them cold-cache, so only the SHARES are comparable with bench.py, not the absolutes.
"""
import argparse
import collections
import csv


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("csv")
    ap.add_argument("--passes", type=int, default=1)
    ap.add_argument("--keep", type=int, default=1)
    ap.add_argument("--out")
    ap.add_argument("--title", default="launch list")
    ap.add_argument("--note", default="")
    ap.add_argument("--ours", action="store_true", help="drop torch's own kernels (input synthesis of the profiled script)")
    a = ap.parse_args()
    rows = list(csv.reader(l for l in open(a.csv) if l.startswith('"')))
    hdr = rows[0]
    ki, vi, ui = hdr.index("Kernel Name"), hdr.index("Metric Value"), hdr.index("Metric Unit")
    if a.ours:
        rows = [hdr] + [r for r in rows[1:] if not any(t in r[ki] for t in ("native::", "at_cuda_detail", "at::", "cub::"))]
    n = len(rows) - 1
    per = n // a.passes
    agg, tot = collections.OrderedDict(), 0.0
    for r in rows[1 + (a.passes - a.keep) * per:]:
        name = r[ki].split("(")[0].replace("void ", "")
        v = float(r[vi].replace(",", ""))
        v = v / 1e3 if r[ui] == "ns" else (v * 1e3 if r[ui] == "ms" else v)
        agg.setdefault(name, [0.0, 0])
        agg[name][0] += v
        agg[name][1] += 1
        tot += v
    lines = [f"# {a.title}", "",
             f"Source: `{a.csv}` (`ncu --metrics gpu__time_duration.sum --clock-control none`), {n} launches = {a.passes} passes of "
             f"{per}; the last {a.keep} averaged.  ncu serialises launches and runs them cold-cache: compare SHARES, not absolutes.",
             "", a.note, "", f"Sum of kernel durations per step: {tot / a.keep:.1f} us", "",
             "| share | us / step | launches / step | kernel |", "|---|---|---|---|"]
    for k, (v, c) in sorted(agg.items(), key=lambda kv: -kv[1][0]):
        lines.append(f"| {v / tot * 100:.2f} % | {v / a.keep:.1f} | {c // a.keep} | `{k}` |")
    text = "\n".join(lines) + "\n"
    if a.out:
        open(a.out, "w").write(text)
    print(text)


if __name__ == "__main__":
    main()
