#!/usr/bin/env python
"""Static summary of the step kernels of one kernel-family object: registers / spills from the ptxas log, SASS instruction mix from
`cuobjdump -sass` (no GPU needed).  Usage: python tools/sass_summary.py [family_real ...]   (default: d3q27_cum_double)"""
import collections
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
OBJ = os.path.join(ROOT, "tnl_lbm_b200", "build")
MODES = {"0": "A-B", "1": "A-A even", "2": "A-A odd"}
GROUPS = [("fp64 DFMA", r"^DFMA"), ("fp64 DADD", r"^DADD"), ("fp64 DMUL", r"^DMUL"), ("fp64 other (MUFU.RCP64H, DSETP ...)", r"^(MUFU|DSETP|DMNMX)"),
          ("fp32 FFMA/FADD/FMUL", r"^(FFMA|FADD|FMUL)"), ("global loads  LDG", r"^LDG"), ("global stores STG", r"^STG"), ("local (spill) LDL/STL", r"^(LDL|STL)"),
          ("integer / address (IMAD, IADD3, LEA, LOP3, SHF, ISETP, SEL ...)", r"^(IMAD|IADD|LEA|LOP3|SHF|ISETP|SEL|IABS|VIADD|VIMNMX|UIADD|ULEA|UIMAD|UMOV|USHF|ULOP|UISETP|USEL)"),
          ("moves / constants (MOV, LDC, S2R ...)", r"^(MOV|LDC|ULDC|S2R|S2UR|CS2R|R2UR|PRMT)"), ("control (BRA, EXIT, BSSY ...)", r"^(BRA|EXIT|BSSY|BSYNC|NOP|CALL|RET|WARPSYNC)")]


def demangle(name):
    return subprocess.run(["c++filt", name], capture_output=True, text=True).stdout.strip()


def main():
    for fam in sys.argv[1:] or ["d3q27_cum_double"]:
        obj = os.path.join(OBJ, f"inst_{fam}.o")
        log = open(obj + ".log").read()
        regs = {m.group(1): (m.group(2), m.group(3), m.group(4)) for m in re.finditer(
            r"Compiling entry function '(\S+)' for 'sm_100a'.*?(\d+) bytes stack frame, (\d+) bytes spill stores.*?Used (\d+) registers", log, re.S)}
        sass = subprocess.run(["cuobjdump", "-sass", obj], capture_output=True, text=True, check=True).stdout
        print(f"== {fam}  ({os.path.relpath(obj, ROOT)}; nvcc -gencode arch=compute_100a,code=sm_100a -O3 -lineinfo)")
        for block in sass.split("Function : ")[1:]:
            name = block.split()[0]
            if "k_bulk" not in name and "k_boundary" not in name:
                continue
            ops = collections.Counter()
            hints = collections.Counter()
            for line in block.splitlines():
                m = re.match(r"\s+/\*[0-9a-f]{4}\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_.]+)", line)
                if not m:
                    continue
                op = m.group(1)
                ops[op.split(".")[0]] += 1
                if op.startswith(("LDG", "STG")):
                    hints[op] += 1
            total = sum(ops.values())
            mode = re.search(r"k_bulkINS_\w+?ELi\d+E[df]Li(\d)E", name)
            stack, spill, nreg = regs.get(name, ("?", "?", "?"))
            title = demangle(name).split("(")[0].replace("void lbmx::", "")
            print(f"\n{title}" + (f"   [{MODES[mode.group(1)]}]" if mode else ""))
            print(f"  registers {nreg}, stack {stack} B, spill stores {spill} B, SASS instructions {total}")
            for label, pat in GROUPS:
                n = sum(c for o, c in ops.items() if re.match(pat, o))
                if n:
                    print(f"  {n:5d}  {label}")
            print("  memory instructions: " + ", ".join(f"{c} x {o}" for o, c in sorted(hints.items())))


if __name__ == "__main__":
    main()
