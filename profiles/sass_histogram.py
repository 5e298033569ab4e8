#!/usr/bin/env python3
"""SASS op-class histogram of selected kernels of the product library (static instruction counts).

usage: sass_histogram.py <lib.so> <kernel substring> [<kernel substring> ...]
Classes: memory (LDS/STS/LDG/STG/LDL/STL/ATOM), integer ALU, video/DPX (VIMNMX3, VIADD..), shuffles/votes, branch / sync,
predicate logic, moves, conversions.
"""
import collections
import re
import subprocess
import sys

CLASSES = [
    ("lds/sts", r"^(LDS|STS|LDSM)"), ("ldg/stg/atom", r"^(LDG|STG|ATOM|RED|LD\b|ST\b|LDC|ULDC)"), ("local", r"^(LDL|STL)"),
    ("shfl/vote/redux", r"^(SHFL|VOTE|VOTEU|REDUX|MATCH)"), ("branch/sync", r"^(BRA|BSSY|BSYNC|BRX|EXIT|WARPSYNC|BAR|CALL|RET|NOP|YIELD|BREAK|BMOV)"),
    ("dpx/video", r"^(VIMNMX|VIADD|VIADDMNMX|VABSDIFF)"), ("imnmx", r"^(IMNMX|UIMNMX)"),
    ("iadd/imad/lea", r"^(IADD|IADD3|IMAD|LEA|UIADD|UIMAD|ULEA|IABS|IDP)"), ("logic/shift", r"^(LOP|LOP3|SHF|SHL|SHR|ULOP|USHF|BFE|BFI|PRMT|UPRMT|SGXT|BREV|POPC|FLO|UFLO|UPOPC)"),
    ("setp/psetp/sel", r"^(ISETP|UISETP|PSETP|PLOP3|UPLOP3|SEL|USEL|FSETP|DSETP|P2R|R2P)"), ("mov", r"^(MOV|UMOV|S2R|S2UR|CS2R|R2UR|UR2R)"),
    ("fp/conv", r"^(F2|I2F|F2I|FADD|FMUL|FFMA|DADD|DMUL|DFMA|MUFU|FMNMX|DMNMX|I2I|F2F)"),
]


def main():
    lib, kerns = sys.argv[1], sys.argv[2:]
    sass = subprocess.run(["cuobjdump", "-sass", lib], capture_output=True, text=True).stdout
    cur, counts = None, {}
    for line in sass.splitlines():
        m = re.search(r"Function : (\S+)", line)
        if m:
            cur = m.group(1)
            continue
        m = re.match(r"\s+/\*[0-9a-f]{4}\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_.]+)", line)
        if m and cur:
            counts.setdefault(cur, collections.Counter())[m.group(1).split(".")[0]] += 1
    for k in kerns:
        for fn, c in counts.items():
            if k not in fn:
                continue
            tot = sum(c.values())
            print(f"## {fn}\n   {tot} instructions")
            left = dict(c)
            for name, rx in CLASSES:
                n = sum(v for op, v in c.items() if re.match(rx, op))
                for op in list(left):
                    if re.match(rx, op):
                        left.pop(op)
                print(f"   {name:18s} {n:6d}  {100.0 * n / tot:5.1f} %")
            other = sum(left.values())
            print(f"   {'other':18s} {other:6d}  {100.0 * other / tot:5.1f} %   {sorted(left.items(), key=lambda kv: -kv[1])[:8]}")


if __name__ == "__main__":
    main()
