"""Per-kernel SASS evidence of the Blackwell paths: counts of the tcgen05 / TMEM / TMA mnemonics in every kernel of
libl2d_b200.so (B200_PROFILING.md: UTCHMMA = tcgen05.mma kind::f16, LDTM / STTM = tcgen05.ld / st, UTMALDG = TMA tensor
load, UTCBAR = tcgen05.commit, UBLKCP = cp.async.bulk, HMMA = legacy mma.sync).  Runs without a GPU.

    python profiles/sass_summary.py > profiles/r2/sass_summary.txt
"""
import collections
import os
import re
import subprocess
import sys

ROOT = os.path.abspath(os.path.join(os.path.dirname(__file__), ".."))
LIB = os.path.join(ROOT, "live2diff_b200", "libl2d_b200.so")
MNEMONICS = ["UTCHMMA", "UTCBAR", "LDTM", "STTM", "UTMALDG", "UTMASTG", "UBLKCP", "HMMA", "LDSM", "SYNCS", "MUFU.EX2", "ACQBULK", "UCGABAR"]


def main():
    sass = subprocess.run(["cuobjdump", "-sass", LIB], capture_output=True, text=True, check=True).stdout
    names = subprocess.run(["cu++filt"], input="\n".join(re.findall(r"Function : (\S+)", sass)), capture_output=True, text=True).stdout.split("\n")
    counts, order, cur, i = {}, [], None, 0
    for line in sass.splitlines():
        m = re.search(r"Function : (\S+)", line)
        if m:
            cur = names[i].strip() if i < len(names) else m.group(1)
            i += 1
            cur = re.sub(r"\(.*", "", cur.replace("(int)", "").replace("(bool)", "").replace("(anonymous namespace)", "<unnamed>"))
            cur = cur.replace("void ", "").replace("l2d::<unnamed>::", "l2d::")
            counts[cur] = collections.Counter()
            order.append(cur)
            continue
        if cur is None:
            continue
        m = re.search(r"^\s+/\*[0-9a-f]+\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_.]+)", line)
        if m:
            op = m.group(1)
            counts[cur]["_total"] += 1
            for mn in MNEMONICS:
                if op.startswith(mn):
                    counts[cur][mn] += 1
    print("kernel".ljust(64) + "".join(m.rjust(9) for m in ["instrs"] + MNEMONICS))
    for k in sorted(order):
        c = counts[k]
        print(k[:63].ljust(64) + str(c["_total"]).rjust(9) + "".join((str(c[m]) if c[m] else ".").rjust(9) for m in MNEMONICS))


if __name__ == "__main__":
    sys.exit(main())
