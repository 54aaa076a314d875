"""Per-kernel summary (and optionally the launch list) of ONE training step from an ncu
`--metrics gpu__time_duration.sum --csv` launch list of bench.py: the launches between the last two optimizer
kernels.  usage: python tools/step_launches.py launches.csv [--list substr,substr...]"""
import collections
import csv
import io
import sys


def main():
    path = sys.argv[1]
    lines = [l for l in open(path) if l.startswith('"')]
    rows = list(csv.DictReader(io.StringIO("".join(lines))))
    names = [r["Kernel Name"].split("(")[0][:90] for r in rows]
    idx = [i for i, n in enumerate(names) if "multi_update" in n or "multi_adam" in n or "multi_sgd" in n]
    a, b = idx[-2] + 1, idx[-1] + 1
    step = rows[a:b]
    agg, tot = collections.OrderedDict(), 0.0
    for r, n in zip(step, names[a:b]):
        ns = float(r["Metric Value"])
        x = agg.setdefault(n, [0, 0.0])
        x[0] += 1
        x[1] += ns
        tot += ns
    print(f"# {path}: one step = {len(step)} launches, {tot / 1e6:.3f} ms (cold-cache, serialised: compare shares)")
    for k, (n, ns) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
        print(f"{ns / 1e3:9.1f} us {100 * ns / tot:5.1f}% {n:3d}  {k}")
    if "--list" in sys.argv:
        keys = sys.argv[sys.argv.index("--list") + 1].split(",")
        for r, n in zip(step, names[a:b]):
            if any(k in n for k in keys):
                print(f"{float(r['Metric Value']) / 1e3:8.1f} us  grid {r['Grid Size']:>16} block {r['Block Size']:>14}  {n[:70]}")


if __name__ == "__main__":
    main()
