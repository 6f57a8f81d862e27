"""Hottest SASS lines of a kernel from `ncu -i X.ncu-rep --page source --csv [--kernel-name ...]` (warp-stall samples)."""
import csv
import sys


def main(path, top=40):
    rows = list(csv.reader(l for l in open(path, errors="replace") if not l.startswith("==")))
    hdr = rows[0]
    idx = {h: i for i, h in enumerate(hdr)}
    samp = next((h for h in hdr if h.startswith("# Samples") or h == "Sampling Data (All)" or "Warp Stall Sampling (All" in h), None)
    src = next((h for h in hdr if h in ("Source", "SASS")), None)
    print("columns:", [h for h in hdr][:14], "| using", samp, src)
    if samp is None or src is None:
        return
    data = []
    for r in rows[1:]:
        try:
            data.append((float(r[idx[samp]].replace(",", "") or 0), r[idx[src]], r[idx.get("Address", 0)] if "Address" in idx else ""))
        except (ValueError, IndexError):
            continue
    tot = sum(d[0] for d in data) or 1.0
    print(f"total samples {tot:.0f}")
    for i, (s, code, addr) in enumerate(data):
        pass
    ranked = sorted(range(len(data)), key=lambda i: -data[i][0])[:top]
    for i in sorted(ranked):
        s, code, addr = data[i]
        print(f"{i:5d} {addr:>8s} {100 * s / tot:6.2f}%  {code[:110]}")


if __name__ == "__main__":
    main(sys.argv[1], int(sys.argv[2]) if len(sys.argv) > 2 else 40)
