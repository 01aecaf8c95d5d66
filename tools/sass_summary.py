#!/usr/bin/env python
"""SASS instruction-class summary of every kernel in libsar.so (nativeness evidence, no GPU needed).

    python tools/sass_summary.py > profiles/r02_sass_summary.csv

Runs `cuobjdump -sass` on the built library and counts, per kernel, the mnemonics that identify the Blackwell paths:
UTCHMMA (tcgen05.mma), LDTM / STTM (tcgen05.ld / st), UTMALDG / UTMASTG (TMA tensor load / store), UBLKCP (cp.async.bulk),
UTCBAR (tcgen05.commit), SYNCS (mbarrier), HMMA (legacy mma.sync), LDGSTS (cp.async), packed fp32 (FFMA2 / FADD2 / FMUL2),
MUFU, plus the total instruction count.
"""
import csv
import re
import subprocess
import sys
from collections import Counter, OrderedDict
from pathlib import Path

LIB = Path(__file__).resolve().parents[1] / "speech_adapter_routing_b200" / "csrc" / "libsar.so"
CLASSES = ["UTCHMMA", "LDTM", "STTM", "UTMALDG", "UTMASTG", "UBLKCP", "UTCBAR", "SYNCS", "HMMA", "LDGSTS", "LDSM",
           "FFMA2", "FADD2", "FMUL2", "MUFU", "LDG", "STG", "LDS", "STS", "SHFL", "BAR"]


def main():
    sass = subprocess.run(["cuobjdump", "-sass", str(LIB)], capture_output=True, text=True, check=True).stdout
    kernels = OrderedDict()
    cur = None
    for line in sass.splitlines():
        m = re.match(r"\s*Function : (\S+)", line)
        if m:
            cur = m.group(1)
            kernels[cur] = Counter()
            continue
        m = re.match(r"\s+/\*[0-9a-f]{4}\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_]+)", line)
        if m and cur:
            op = m.group(1)
            kernels[cur]["total"] += 1
            for c in CLASSES:
                if op == c or op.startswith(c + ".") or (c in ("UTCHMMA", "UTMALDG", "UTMASTG", "UBLKCP", "UTCBAR", "LDTM", "STTM") and op.startswith(c)):
                    kernels[cur][c] += 1
    demangle = subprocess.run(["c++filt"], input="\n".join(kernels), capture_output=True, text=True).stdout.splitlines()
    w = csv.writer(sys.stdout)
    w.writerow(["kernel", "total"] + CLASSES)
    for (name, cnt), dm in zip(kernels.items(), demangle):
        short = re.sub(r"\(.*", "", dm).replace("void ", "")
        w.writerow([short, cnt["total"]] + [cnt[c] for c in CLASSES])


if __name__ == "__main__":
    main()
