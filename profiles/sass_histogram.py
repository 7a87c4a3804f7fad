#!/usr/bin/env python
"""Opcode histogram of one kernel's SASS (cuobjdump -sass on the built library): the evidence for tcgen05 / TMA / packed-FMA use
and for register spills (STL / LDL).   python profiles/sass_histogram.py k_tc16 [more kernels] > profiles/r2_sass_<kernel>.txt"""
import collections
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIB = os.path.join(ROOT, "idff_b200", "lib", "libidff_b200.so")
KEY = ("UTCHMMA", "UTCQMMA", "UTCBAR", "LDTM", "STTM", "UBLKCP", "UTMALDG", "SYNCS", "FFMA2", "FFMA", "FMUL2", "FADD2", "HMMA", "MUFU",
       "STAS", "STL", "LDL", "BSSY", "BSYNC", "STS", "LDS", "LDG", "STG", "ELECT", "R2UR", "F2FP", "BAR", "ATOM", "RED")


def main():
    sass = subprocess.run(["cuobjdump", "-sass", LIB], capture_output=True, text=True, check=True).stdout
    res = subprocess.run(["cuobjdump", "-res-usage", LIB], capture_output=True, text=True).stdout
    for want in sys.argv[1:]:
        blocks = re.split(r"\n\s*Function : ", sass)
        for blk in blocks[1:]:
            name = blk.split("\n", 1)[0].strip()
            if want not in name:
                continue
            ops = collections.Counter()
            n = 0
            for line in blk.splitlines():
                m = re.match(r"\s*/\*[0-9a-f]{4,}\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_.]+)", line)
                if m:
                    ops[m.group(1)] += 1
                    n += 1
            fam = collections.Counter()
            for op, c in ops.items():
                fam[op.split(".")[0]] += c
            print(f"==== {name}: {n} SASS instructions")
            for line in res.splitlines():
                if name in line:
                    print("  res-usage:", line.split(name)[-1].strip(" :"))
            print("  key families: " + ", ".join(f"{k} {fam[k]}" for k in KEY if fam.get(k)))
            print("  full mnemonics (count >= 2 or tensor/TMA/spill related):")
            for op, c in sorted(ops.items(), key=lambda kv: -kv[1]):
                if c >= 2 or any(op.startswith(k) for k in KEY[:12]):
                    print(f"    {c:6d}  {op}")
            print()


if __name__ == "__main__":
    main()
