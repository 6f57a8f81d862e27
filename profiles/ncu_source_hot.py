"""Hottest SASS lines of each kernel in `ncu -i X.ncu-rep --page source --csv` output (warp-stall samples per instruction,
with the dominant stall reasons and the execution count).  usage: ncu_source_hot.py file.csv [top] [kernel-substring]"""
import csv
import sys


def kernels(path):
    rows = list(csv.reader(l for l in open(path, errors="replace") if not l.startswith("==")))
    cur, name, hdr = None, None, None
    for r in rows:
        if r and r[0] == "Kernel Name":
            if cur:
                yield name, hdr, cur
            name, hdr, cur = r[1], None, []
        elif r and r[0] == "Address":
            hdr = r
        elif cur is not None and hdr is not None and r:
            cur.append(r)
    if cur:
        yield name, hdr, cur


def main(path, top=40, only=""):
    for name, hdr, rows in kernels(path):
        if only and only not in name:
            continue
        idx = {h: i for i, h in enumerate(hdr)}
        si, ci, ei = idx["# Samples"], idx["Source"], idx["Instructions Executed"]
        stalls = [h for h in hdr if h.startswith("stall_") and "(Not Issued)" not in h]
        data = []
        for r in rows:
            try:
                s = float(r[si] or 0)
            except (ValueError, IndexError):
                continue
            st = sorted(((float(r[idx[h]] or 0), h[6:]) for h in stalls), reverse=True)[:3]
            data.append((s, r[ci].strip(), r[ei], " ".join(f"{n}:{int(v)}" for v, n in st if v > 0)))
        tot = sum(d[0] for d in data) or 1.0
        print(f"=== {name[:120]}\n    {len(data)} instructions, {tot:.0f} samples")
        agg = {}
        for r in rows:
            for h in stalls:
                try:
                    agg[h[6:]] = agg.get(h[6:], 0.0) + float(r[idx[h]] or 0)
                except (ValueError, IndexError):
                    pass
        print("    stall totals:", " ".join(f"{k}:{100 * v / tot:.1f}%" for k, v in sorted(agg.items(), key=lambda kv: -kv[1])[:9]))
        ranked = sorted(range(len(data)), key=lambda i: -data[i][0])[:top]
        for i in sorted(ranked):
            s, code, ex, st = data[i]
            print(f"{i:5d} {100 * s / tot:6.2f}%  x{ex:>8s}  {code[:64]:64s} {st}")


if __name__ == "__main__":
    main(sys.argv[1], int(sys.argv[2]) if len(sys.argv) > 2 else 40, sys.argv[3] if len(sys.argv) > 3 else "")
