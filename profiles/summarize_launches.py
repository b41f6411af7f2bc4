"""Summarise an `ncu --metrics gpu__time_duration.sum --csv` launch list: one training step
(the launches between the last two clip_adam_kernel launches), aggregated per kernel.

    python profiles/summarize_launches.py gpurun_out/launches.csv > profiles/<name>.txt
"""
import collections
import csv
import re
import sys


def main(path):
    with open(path) as f:
        lines = [l for l in f if not l.startswith("==")]
    rows = list(csv.DictReader(lines))
    names = [re.sub(r"\(.*", "", r["Kernel Name"]) for r in rows]
    adam = [i for i, n in enumerate(names) if "clip_adam" in n]
    a, b = adam[-2] + 1, adam[-1] + 1
    agg, tot = collections.OrderedDict(), 0.0
    for row, n in zip(rows[a:b], names[a:b]):
        t = float(row["Metric Value"].replace(",", ""))
        t = {"ns": t / 1e3, "us": t, "ms": t * 1e3, "s": t * 1e6}.get(row["Metric Unit"], t)
        c = agg.setdefault(n[:90] + "  " + row["Grid Size"].replace(" ", ""), [0, 0.0, "", row["Block Size"]])
        c[0] += 1
        c[1] += t
        tot += t
    print(f"# {path}: one train step = {b - a} launches, {tot:.1f} us summed kernel time "
          f"(ncu: cold caches, serialised -- compare shares, not absolutes)")
    print(f"{'us':>10} {'share':>6} {'count':>5}  kernel  [grid] [block]")
    for k, (c, t, g, bl) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
        print(f"{t:10.1f} {100 * t / tot:5.1f}% {c:5d}  {k}  {g} {bl}")


if __name__ == "__main__":
    main(sys.argv[1])
