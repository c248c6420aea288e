#!/usr/bin/env python3
"""Per-loop opcode / pipe histogram of one kernel from `cuobjdump -sass` (stdin or file): loops = backward branches.
usage: cuobjdump -sass x.cubin | tools/loop_hist.py [kernel-substring]"""
import re
import sys
from collections import Counter

ALU = {"IADD3", "LOP3", "SHF", "LEA", "ISETP", "SEL", "PRMT", "VIADD", "IABS", "FLO", "POPC", "MOV", "PLOP3", "VIMNMX", "IMNMX", "BMSK", "SGXT"}
FP64 = {"DADD", "DFMA", "DMUL", "DSETP"}


def main():
    pat = sys.argv[1] if len(sys.argv) > 1 else ""
    on, ins = False, []
    for line in sys.stdin:
        if "Function :" in line:
            on = pat in line
            continue
        if not on:
            continue
        m = re.match(r"\s+/\*([0-9a-f]{4,6})\*/\s+(.*?);", line)
        if m:
            t = re.sub(r"^@!?U?P\d\s+", "", m.group(2).strip())
            ins.append((int(m.group(1), 16), t))
    loops = []
    for a, t in ins:
        m = re.match(r"BRA(?:\.U)?\s+(?:!?U?P\d,\s+)?(0x[0-9a-f]+)", t)
        if m and int(m.group(1), 16) < a:
            loops.append((int(m.group(1), 16), a))
    print(f"kernel instructions: {len(ins)} ({16 * len(ins) / 1024:.1f} KiB)")
    for lo, hi in sorted(loops, key=lambda x: x[0] - x[1])[:int(sys.argv[2]) if len(sys.argv) > 2 else 4]:
        c = Counter(t.split()[0] for a, t in ins if lo <= a <= hi)
        base = lambda k: k.split(".")[0]
        alu = sum(v for k, v in c.items() if base(k) in ALU)
        wide = sum(v for k, v in c.items() if k.startswith("IMAD.WIDE"))
        imad = sum(v for k, v in c.items() if base(k) == "IMAD")
        f64 = sum(v for k, v in c.items() if base(k) in FP64)
        tot = sum(c.values())
        print(f"loop {lo:#x}..{hi:#x}: {tot} instructions | alu {alu} | fmaheavy {imad} instr = {imad + wide} slots (IMAD.WIDE {wide}) | fp64 {f64} | issue slots {tot + f64}")
        print("    " + ", ".join(f"{k}:{v}" for k, v in c.most_common(12)))


if __name__ == "__main__":
    main()
