"""Condenses an `ncu --page source --csv` dump into runs of SASS with equal execution counts.

    ncu -i rep.ncu-rep --page source --csv --kernel-name regex:NAME > src.csv
    python tools/sass_segments.py src.csv [min_share_percent]
"""
import csv
import sys


def main():
    rows = list(csv.reader(open(sys.argv[1])))
    floor = float(sys.argv[2]) if len(sys.argv) > 2 else 0.4
    hdr = next(r for r in rows if "Instructions Executed" in r)
    ia, ie, isamp = hdr.index("Source"), hdr.index("Instructions Executed"), hdr.index("# Samples")
    data = [r for r in rows if len(r) == len(hdr) and r[ie].isdigit()]
    tot = sum(int(r[ie]) for r in data)
    print("total warp instructions", tot, "sass lines", len(data))
    segs, prev, start = [], None, 0
    for k, r in enumerate(data):
        e = int(r[ie])
        if prev is None or abs(e - prev) > 0.02 * max(prev, 1):
            if prev is not None:
                segs.append((start, k - 1, prev))
            start, prev = k, e
    segs.append((start, len(data) - 1, prev))
    for a, b, e in segs:
        n = b - a + 1
        samp = sum(int(data[i][isamp]) for i in range(a, b + 1))
        ops = {}
        for i in range(a, b + 1):
            t = data[i][ia].split()
            op = (t[1] if t[0].startswith("@") else t[0]).split(".")[0]
            ops[op] = ops.get(op, 0) + 1
        top = sorted(ops.items(), key=lambda x: -x[1])[:7]
        if e * n > tot * floor / 100:
            print(f"sass {a:5d}-{b:5d} n={n:4d} exec/each={e:9d} share={e * n / tot * 100:5.1f}% samples={samp:6d} {top}")


if __name__ == "__main__":
    main()
