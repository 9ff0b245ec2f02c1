"""Per-kernel SASS opcode histogram of libmeshgen_b200.so (cuobjdump -sass): instruction count, image size, the opcodes
that prove the Blackwell-native path (UBLKCP = cp.async.bulk, SYNCS = mbarrier, FP64 pipe) and the local-memory
instructions (LDL / STL = spills).  Usage: python profiles/sass_histogram.py [lib.so] > profiles/r2_sass_histogram.txt"""
import re
import subprocess
import sys
from collections import Counter

lib = sys.argv[1] if len(sys.argv) > 1 else "reinforcementlearning4meshgeneration_b200/lib/libmeshgen_b200.so"
txt = subprocess.run(["cuobjdump", "-sass", lib], capture_output=True, text=True).stdout
arch = sorted(set(re.findall(r"arch = (sm_\w+)", txt)))
print(f"{lib}: arch {arch}")
for fn in re.split(r"\n\s*Function : ", txt)[1:]:
    name = subprocess.run(["c++filt", fn.split("\n")[0].strip()], capture_output=True, text=True).stdout.strip().split("(")[0]
    ops = Counter()
    for m in re.finditer(r"/\*[0-9a-f]{4,}\*/\s+(?:@!?U?P\d+\s+)?([A-Z][A-Z0-9_]*)", fn):
        ops[m.group(1)] += 1
    n = sum(ops.values())
    fp64 = sum(v for k, v in ops.items() if k in ("DFMA", "DMUL", "DADD", "DSETP", "DMNMX"))
    print(f"\n{name}: {n} instructions ({n * 16 // 1024} KB)  FP64 {fp64}  UBLKCP {ops['UBLKCP']}  SYNCS {ops['SYNCS']}  "
          f"LDL {ops['LDL']}  STL {ops['STL']}  ATOM/ATOMG/RED {ops['ATOM'] + ops['ATOMG'] + ops['RED'] + ops['REDG']}  "
          f"SHFL {ops['SHFL']}  VOTE {ops['VOTE']}  tensor-core (HMMA/UTC*MMA) {sum(v for k, v in ops.items() if 'MMA' in k)}")
    print("   " + "  ".join(f"{k} {v}" for k, v in ops.most_common(18)))
