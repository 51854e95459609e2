#!/usr/bin/env python
"""tools/sass_histogram.py — static opcode histogram of the solve kernel's SASS (cuobjdump -sass of the shipped
library), written next to the ncu executed-opcode histogram of profiles/r02_ncu_summary.md.  No GPU needed."""
import collections
import re
import subprocess
import sys

LIB = "mpc_trajectory_generator_b200/libnmpc_b200.so"
KERNEL = sys.argv[1] if len(sys.argv) > 1 else "_Z17nmpc_solve_kernelILi8ELi3ELb0EEv5KArgs"   # N = 20: G = 8, S = 3, without helper warps (Lb1: with)


def main():
    txt = subprocess.run(["cuobjdump", "-sass", LIB], capture_output=True, text=True, check=True).stdout
    m = re.search(r"Function : " + re.escape(KERNEL) + r"\n(.*?)(?=\n\s*Function : |\Z)", txt, re.S)
    ops = collections.Counter()
    total = 0
    for line in m.group(1).splitlines():
        mm = re.match(r"\s+/\*[0-9a-f]{4,}\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_]+(?:\.[A-Z0-9_.]+)?)", line)
        if mm:
            ops[mm.group(1).split(".")[0]] += 1
            total += 1
    res = subprocess.run(["cuobjdump", "-res-usage", LIB], capture_output=True, text=True).stdout
    usage = ""
    lines = res.splitlines()
    for i, l in enumerate(lines):
        if KERNEL in l and i + 1 < len(lines):
            usage = lines[i + 1].strip()
    print(f"# static SASS opcode histogram — {KERNEL}\n")
    print(f"`cuobjdump -sass {LIB}`; {total} instructions = {total * 16 / 1024:.1f} KB of code; `{usage}`\n")
    fam = {"FP64": ("DFMA", "DADD", "DMUL", "DSETP", "DMNMX"), "shuffle / vote / redux": ("SHFL", "VOTE", "REDUX", "MATCH"),
           "shared memory": ("LDS", "STS", "LDSM"), "global / local memory": ("LDG", "STG", "LDL", "STL", "LD", "ST", "ATOMG", "RED"),
           "barriers": ("WARPSYNC", "BAR", "BSYNC", "BSSY", "NANOSLEEP"), "branches": ("BRA", "BRX", "EXIT", "CALL", "RET", "JMP")}
    print("| family | static count | share |\n|---|---|---|")
    for name, keys in fam.items():
        n = sum(ops[k] for k in keys)
        print(f"| {name} | {n} | {100 * n / total:.1f} % |")
    print("\n| opcode | static count | share |\n|---|---|---|")
    for op, n in ops.most_common(40):
        print(f"| {op} | {n} | {100 * n / total:.1f} % |")


if __name__ == "__main__":
    main()
