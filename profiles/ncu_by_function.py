"""Attribute ncu warp-stall samples / executed instructions to the device functions of a kernel.
Usage: python profiles/ncu_by_function.py <lib.so> <source_page.csv> <kernel substring>
(source page: ncu -i X.ncu-rep --page source --csv --kernel-name regex:<k>)"""
import csv
import os
import re
import subprocess
import sys
import tempfile
from collections import Counter


def symbols(lib, kernel):
    d = tempfile.mkdtemp()
    subprocess.run(["cuobjdump", "-xelf", "all", os.path.abspath(lib)], cwd=d, capture_output=True)
    cubin = [f for f in os.listdir(d) if f.endswith(".cubin")][0]
    out = subprocess.run(["readelf", "-sW", os.path.join(d, cubin)], capture_output=True, text=True).stdout
    syms = []
    for line in out.splitlines():
        f = line.split()
        if len(f) < 8 or f[3] != "FUNC" or kernel not in f[-1]:
            continue
        name = f[-1]
        short = name.split("$")[-1] if "$" in name else "<kernel body>"
        short = re.sub(r"^_ZN2mg\d+", "", short)
        short = re.sub(r"E(RK|NS|dd|Ed|dP).*$", "", short)
        size = int(f[2], 16) if f[2].startswith("0x") else int(f[2])
        syms.append((int(f[1], 16), size, short))
    return sorted(syms)


def main(lib, src, kernel):
    syms = symbols(lib, kernel)
    rows = list(csv.reader(open(src)))
    hdr = next(r for r in rows if r and r[0] == "Address")
    col = {h: i for i, h in enumerate(hdr)}
    data, seen = [], set()
    for r in rows:
        if len(r) == len(hdr) and r[0].startswith("0x") and r[0] not in seen:
            seen.add(r[0])
            data.append(r)
    base = min(int(r[0], 16) for r in data)
    lo = min(s[0] for s in syms)
    smp, ex, stall = Counter(), Counter(), {}
    for r in data:
        off = int(r[0], 16) - base + lo
        name = min(((sz, n) for s0, sz, n in syms if s0 <= off < s0 + sz), default=(0, "?"))[1]
        smp[name] += int(r[col["# Samples"]] or 0)
        ex[name] += int(r[col["Instructions Executed"]] or 0)
        st = stall.setdefault(name, Counter())
        for k in ("stall_long_sb", "stall_wait", "stall_no_inst", "stall_short_sb", "stall_branch_resolving", "stall_math"):
            st[k] += int(r[col[k]] or 0)
    ts, te = sum(smp.values()), sum(ex.values())
    print(f"kernel {kernel}: {ts} samples, {te} warp instructions, {len(data)} SASS instructions")
    for n, c in smp.most_common(20):
        st = stall[n]
        top = ", ".join(f"{k[6:]} {100*v/max(1,c):.0f}%" for k, v in st.most_common(3))
        print(f"  {n:32s} samples {100*c/ts:5.1f}%  instr {100*ex[n]/te:5.1f}%   [{top}]")


if __name__ == "__main__":
    main(sys.argv[1], sys.argv[2], sys.argv[3])
