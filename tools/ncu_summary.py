"""Summarises an `ncu --metrics gpu__time_duration.sum --csv` launch list into a per-kernel table.

    python tools/ncu_summary.py gpurun_out/rX/launches.csv [--first N] [--last N] > profiles/rX_launch_summary.md
"""
import argparse
import collections
import csv
import re


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("csv")
    ap.add_argument("--first", type=int, default=0)
    ap.add_argument("--last", type=int, default=1 << 30)
    a = ap.parse_args()
    rows = []
    with open(a.csv, newline="") as f:
        lines = [ln for ln in f if ln.startswith('"')]
    rd = csv.reader(lines)
    hdr = next(rd)
    i_name, i_val, i_unit = hdr.index("Kernel Name"), hdr.index("Metric Value"), hdr.index("Metric Unit")
    for n, r in enumerate(rd):
        if n < a.first or n >= a.last:
            continue
        v = float(r[i_val].replace(",", ""))
        if r[i_unit] in ("us", "usecond"):
            v *= 1e3
        elif r[i_unit] in ("ms", "msecond"):
            v *= 1e6
        rows.append((r[i_name], v))
    tot = sum(v for _, v in rows)
    agg = collections.OrderedDict()
    for name, v in rows:
        short = re.sub(r"\(.*", "", name)
        short = short if len(short) <= 80 else short[:80]
        c = agg.setdefault(short, [0, 0.0])
        c[0] += 1
        c[1] += v
    print(f"total {tot / 1e6:.1f} ms over {len(rows)} launches\n")
    print("| kernel | launches | total us | share | avg us |")
    print("|---|---|---|---|---|")
    for name, (n, v) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
        if v / tot < 0.002:
            continue
        print(f"| `{name}` | {n} | {v / 1e3:.0f} | {100 * v / tot:.1f}% | {v / 1e3 / n:.1f} |")


if __name__ == "__main__":
    main()
