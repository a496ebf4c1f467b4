"""Aggregate an `ncu --metrics gpu__time_duration.sum --csv` launch list by kernel name."""
import collections
import csv
import sys


def main(path, last=None):
    with open(path) as f:
        lines = [l for l in f if not l.startswith("==")]
    rows = list(csv.DictReader(lines))
    if last:
        rows = rows[-int(last):]
    agg = collections.defaultdict(lambda: [0, 0.0])
    for x in rows:
        name = x["Kernel Name"]
        v = float(x["Metric Value"].replace(",", ""))
        u = x["Metric Unit"]
        v = v / 1000 if u == "ns" else (v * 1000 if u == "ms" else v)
        key = name[:70] + " grid=" + x["Grid Size"].replace(" ", "") if "--grid" in sys.argv else name[:90]
        agg[key][0] += 1
        agg[key][1] += v
    tot = sum(v[1] for v in agg.values())
    for k, v in sorted(agg.items(), key=lambda kv: -kv[1][1]):
        print(f"{k:100s} n={v[0]:5d} total={v[1]:10.1f} us avg={v[1] / v[0]:9.2f} us share={v[1] / tot:.3f}")
    print("total us", round(tot, 1), "launches", len(rows))


if __name__ == "__main__":
    args = [a for a in sys.argv[1:] if not a.startswith("--")]
    main(*args)
