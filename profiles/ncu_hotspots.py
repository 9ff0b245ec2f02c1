"""Aggregate an `ncu --page source --csv` export: top SASS instructions by warp-stall samples and
samples per 2 KB code bucket.  Usage: python profiles/ncu_hotspots.py src.csv [top]"""
import csv
import sys
from collections import Counter


def main(path, ntop=40):
    rows = list(csv.reader(open(path)))
    hdr = next(r for r in rows if r and r[0] == "Address")
    data = [r for r in rows if len(r) == len(hdr) and r[0].startswith("0x")]
    col = {h: i for i, h in enumerate(hdr)}
    g = lambda r, k: int(r[col[k]] or 0)
    tot = sum(g(r, "# Samples") for r in data)
    base = int(data[0][0], 16)
    print("total samples", tot, "instructions", len(data))
    for r in sorted(data, key=lambda r: -g(r, "# Samples"))[:ntop]:
        print(f"{int(r[0],16)-base:7x} {g(r,'# Samples'):6d} {100*g(r,'# Samples')/tot:5.1f}% ex={g(r,'Instructions Executed'):>9d} "
              f"long={g(r,'stall_long_sb'):>5d} wait={g(r,'stall_wait'):>5d} noinst={g(r,'stall_no_inst'):>5d} "
              f"short={g(r,'stall_short_sb'):>5d} br={g(r,'stall_branch_resolving'):>4d}  {r[col['Source']].strip()[:64]}")
    c, ex = Counter(), Counter()
    for r in data:
        k = (int(r[0], 16) - base) // 0x800
        c[k] += g(r, "# Samples")
        ex[k] += g(r, "Instructions Executed")
    tex = sum(ex.values())
    for k in sorted(c):
        if c[k] > tot * 0.01:
            print(f"bucket {k*0x800:6x}: {100*c[k]/tot:5.1f}% samples, {100*ex[k]/tex:5.1f}% of executed instructions")


if __name__ == "__main__":
    main(sys.argv[1], int(sys.argv[2]) if len(sys.argv) > 2 else 40)
