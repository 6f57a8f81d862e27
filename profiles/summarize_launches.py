"""Summarise an ncu `gpu__time_duration.sum` launch list (CSV) per kernel: count, total, average, share."""
import collections
import csv
import re
import sys


def load(path):
    lines = [l for l in open(path) if not l.startswith("==")]
    rows = []
    for r in csv.DictReader(lines):
        v = float(r["Metric Value"].replace(",", ""))
        unit = r["Metric Unit"]
        ns = v * 1e3 if unit in ("us", "usecond") else v if unit in ("ns", "nsecond") else v * 1e6
        rows.append((re.sub(r"\(.*", "", r["Kernel Name"]).replace("void ", "").replace("l2d::", ""), ns, r["Grid Size"]))
    return rows


def main(path, top=0):
    rows = load(path)
    agg = collections.defaultdict(lambda: [0, 0.0])
    for name, ns, _ in rows:
        agg[name][0] += 1
        agg[name][1] += ns
    tot = sum(v[1] for v in agg.values())
    print(f"{len(rows)} launches, {tot / 1e6:.3f} ms summed kernel time (ncu: serialised, cold caches -> compare SHARES)")
    print(f"{'kernel':52s} {'n':>4s} {'total ms':>9s} {'avg us':>8s} {'share':>6s}")
    for k, v in sorted(agg.items(), key=lambda kv: -kv[1][1]):
        print(f"{k[:52]:52s} {v[0]:4d} {v[1] / 1e6:9.3f} {v[1] / v[0] / 1e3:8.1f} {v[1] / tot * 100:5.1f}%")
    if top:
        print(f"\ntop {top} single launches:")
        for name, ns, grid in sorted(rows, key=lambda r: -r[1])[:top]:
            print(f"  {ns / 1e3:8.1f} us  {name[:44]:44s} grid {grid}")


if __name__ == "__main__":
    main(sys.argv[1], int(sys.argv[2]) if len(sys.argv) > 2 else 0)
