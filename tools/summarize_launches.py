"""Summarise an ncu `--metrics gpu__time_duration.sum --csv` launch list per kernel.
usage: python tools/summarize_launches.py launches.csv [skip_first_n]"""
import collections
import csv
import io
import sys


def main():
    path = sys.argv[1]
    skip = int(sys.argv[2]) if len(sys.argv) > 2 else 0
    lines = [l for l in open(path) if l.startswith('"')]
    rows = list(csv.DictReader(io.StringIO("".join(lines))))[skip:]
    agg, tot = collections.OrderedDict(), 0.0
    for row in rows:
        k, ns = row["Kernel Name"], float(row["Metric Value"])
        a = agg.setdefault(k, [0, 0.0])
        a[0] += 1
        a[1] += ns
        tot += ns
    print(f"# {path}: {len(rows)} launches, {tot / 1e6:.3f} ms total (cold-cache, serialised: compare shares)")
    print(f"# {'ms':>9} {'share':>6} {'count':>5}  kernel")
    for k, (n, ns) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
        print(f"{ns / 1e6:11.3f} {100 * ns / tot:5.1f}% {n:5d}  {k[:150]}")


if __name__ == "__main__":
    main()
