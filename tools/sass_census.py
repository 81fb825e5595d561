#!/usr/bin/env python
"""Per-kernel SASS opcode census of a cubin / .so (cuobjdump -sass): the evidence table for profiles/.

    python tools/sass_census.py audio_b200/lib/libb200audio.so [regex]
Counts the mnemonics that matter on sm_100a: UTCHMMA (tcgen05.mma), LDTM (tcgen05.ld), UBLKCP / UTMALDG (TMA),
FFMA2 / FADD2 / FMUL2 (packed FP32), HMMA (legacy mma.sync), plus FFMA / FADD / FMUL / MOV / SHFL / LDS / STS totals.
"""
import collections
import re
import subprocess
import sys

KEYS = ["UTCHMMA", "LDTM", "UTCBAR", "UBLKCP", "UTMALDG", "FFMA2", "FADD2", "FMUL2", "HMMA", "FFMA", "FADD", "FMUL",
        "MOV", "SHFL", "LDS", "STS", "LDG", "STG", "F2FP", "SYNCS", "total"]


def main():
    path = sys.argv[1]
    pat = re.compile(sys.argv[2]) if len(sys.argv) > 2 else None
    out = subprocess.run(["cuobjdump", "-sass", path], capture_output=True, text=True).stdout
    name, counts = None, collections.OrderedDict()
    for line in out.splitlines():
        m = re.match(r"\s*Function : (\S+)", line)
        if m:
            name = m.group(1)
            counts[name] = collections.Counter()
            continue
        m = re.match(r"\s*/\*[0-9a-f]{4,}\*/\s+(?:@!?U?P\d\s+)?([A-Z0-9_]+)", line)
        if m and name:
            op = m.group(1)
            counts[name][op] += 1
            counts[name]["total"] += 1
    print("kernel," + ",".join(KEYS))
    for name, c in counts.items():
        if pat and not pat.search(name):
            continue
        demangled = subprocess.run(["c++filt", name], capture_output=True, text=True).stdout.strip()
        short = re.sub(r"\(.*", "", demangled.replace("(anonymous namespace)::", "")).replace("b200a::", "")
        print(short + "," + ",".join(str(c.get(k, 0)) for k in KEYS))


if __name__ == "__main__":
    main()
