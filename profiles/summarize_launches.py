"""Summarise an `ncu --metrics gpu__time_duration.sum --csv` launch list: per-kernel totals and
the launch sequence.  Usage: python profiles/summarize_launches.py gpurun_out/launches.csv [--seq N]"""
import collections
import csv
import re
import sys


def load(path):
    lines = [l for l in open(path) if l.startswith('"')]
    out = []
    for row in csv.DictReader(lines):
        v = float(row["Metric Value"].replace(",", ""))
        unit = row["Metric Unit"]
        v = v / 1000 if unit in ("ns", "nsecond") else (v * 1000 if unit in ("ms", "msecond") else v)
        name = row["Kernel Name"]
        short = re.sub(r"^void ", "", name).replace("<unnamed>::", "").replace("(anonymous namespace)::", "")
        short = re.sub(r"\(.*", "", short)
        out.append((short, v, row.get("Grid Size", ""), row.get("Block Size", "")))
    return out


def main():
    rows = load(sys.argv[1])
    agg = collections.OrderedDict()
    for name, us, *_ in rows:
        base = re.sub(r"<.*", "", name)
        a = agg.setdefault(base, [0, 0.0])
        a[0] += 1
        a[1] += us
    total = sum(t for _, t in agg.values())
    print(f"{len(rows)} launches, {total:.1f} us total (cold-cache, serialised: compare shares)")
    for k, (c, t) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
        print(f"{t:11.1f} us {100 * t / total:5.1f}%  n={c:4d}  avg={t / c:9.1f} us  {k}")
    if "--seq" in sys.argv:
        n = int(sys.argv[sys.argv.index("--seq") + 1])
        print("---- launch sequence")
        for name, us, grid, block in rows[:n]:
            print(f"{us:10.1f} us  {grid:>14} {block:>12}  {name[:110]}")


if __name__ == "__main__":
    main()
