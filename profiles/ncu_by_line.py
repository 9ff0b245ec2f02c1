"""Attribute ncu warp-stall samples / executed instructions of a kernel to CUDA source lines by joining the SASS
source page with nvdisasm's line table (innermost inlined location).
Usage: python profiles/ncu_by_line.py <lib.so> <source_page.csv> <kernel substring> [top]
(source page: ncu -i X.ncu-rep --page source --csv --kernel-name regex:<k>)"""
import csv
import os
import re
import subprocess
import sys
import tempfile
from collections import defaultdict


def line_table(lib, kernel):
    d = tempfile.mkdtemp()
    subprocess.run(["cuobjdump", "-xelf", "all", os.path.abspath(lib)], cwd=d, capture_output=True)
    cubin = [f for f in os.listdir(d) if f.endswith(".cubin")][0]
    txt = subprocess.run(["nvdisasm", "-g", "-c", os.path.join(d, cubin)], capture_output=True, text=True).stdout
    table, inside, cur = {}, False, None
    for ln in txt.splitlines():
        if ln.startswith("\t.section\t.text."):
            inside = kernel in ln and "_kernel" in ln
            continue
        if not inside:
            continue
        m = re.search(r'//## File "([^"]+)", line (\d+)', ln)
        if m:
            cur = (os.path.basename(m.group(1)), int(m.group(2)))
            continue
        m = re.match(r"\s*/\*([0-9a-f]{4,})\*/", ln)
        if m:
            table[int(m.group(1), 16)] = cur
    return table


def main(lib, path, kernel, ntop=40):
    table = line_table(lib, kernel)
    rows = list(csv.reader(open(path)))
    hdr = next(r for r in rows if r and r[0] == "Address")
    col = {h: i for i, h in enumerate(hdr)}
    data = [r for r in rows if len(r) == len(hdr) and r[0].startswith("0x")]
    base = int(data[0][0], 16)
    seen = set()
    smp, ex, st = defaultdict(int), defaultdict(int), defaultdict(lambda: defaultdict(int))
    for r in data:
        off = int(r[0], 16) - base
        if off in seen:              # the page repeats per captured launch
            continue
        seen.add(off)
        key = table.get(off)
        smp[key] += int(r[col["# Samples"]] or 0)
        ex[key] += int(r[col["Instructions Executed"]] or 0)
        for k in ("stall_long_sb", "stall_wait", "stall_no_inst", "stall_short_sb", "stall_branch_resolving"):
            st[key][k[6:]] += int(r[col[k]] or 0)
    ts, te = sum(smp.values()), sum(ex.values())
    src = {}
    print(f"kernel {kernel}: {ts} samples, {te} warp instructions")
    for key in sorted(smp, key=lambda k: -smp[k])[:ntop]:
        text = ""
        if key:
            f = os.path.join(os.path.dirname(os.path.abspath(lib)), "..", "csrc", key[0])
            if os.path.exists(f):
                src.setdefault(f, open(f).read().splitlines())
                text = src[f][key[1] - 1].strip()[:70] if key[1] - 1 < len(src[f]) else ""
        top = sorted(st[key].items(), key=lambda kv: -kv[1])[:2]
        why = ", ".join(f"{k} {100*v/max(1,smp[key]):.0f}%" for k, v in top)
        print(f"  {str(key[0])+':'+str(key[1]) if key else '?':22s} smp {100*smp[key]/ts:5.1f}%  instr {100*ex[key]/te:5.1f}%  [{why}]  {text}")


if __name__ == "__main__":
    main(sys.argv[1], sys.argv[2], sys.argv[3], int(sys.argv[4]) if len(sys.argv) > 4 else 40)
