#!/usr/bin/env python
"""tools/sass_cycles.py — static single-warp issue estimate of a kernel's loops from its SASS (no GPU needed).

    python tools/sass_cycles.py <lib.so> <kernel-substring> [--dump lo hi]

Every sm_100 instruction carries its issue stall count in bits [105:109) of its encoding (the cycles the warp
waits before its next instruction; fixed-latency dependencies — DFMA/DADD/DMUL/FSEL chains — are encoded there by
ptxas).  Summing the field over a loop body gives the cycles a LONE warp spends issuing it, a lower bound that
ignores scoreboard waits (LDS, SHFL) — the figure that bounds the tail of a batch.  Prints every backward branch
(= loop) with its body length, stall sum and opcode mix, innermost first."""
import collections
import re
import subprocess
import sys


def sass(lib, kernel):
    out = subprocess.run(["cuobjdump", "-sass", lib], capture_output=True, text=True).stdout
    fn, cur = None, []
    res = {}
    for line in out.splitlines():
        m = re.search(r"Function : (\S+)", line)
        if m:
            fn = m.group(1)
            res[fn] = []
            continue
        m = re.match(r"\s*/\*([0-9a-f]{4,})\*/\s+(.*?);\s*/\* 0x([0-9a-f]{16}) \*/", line)
        if m and fn:
            res[fn].append([int(m.group(1), 16), m.group(2).strip(), int(m.group(3), 16), None])
            continue
        m = re.match(r"\s*/\* 0x([0-9a-f]{16}) \*/", line)
        if m and fn and res[fn]:
            res[fn][-1][3] = int(m.group(1), 16)
    for k, v in res.items():
        if kernel in k:
            return k, v
    raise SystemExit(f"no kernel matching {kernel}: {list(res)[:20]}")


def main():
    lib, kernel = sys.argv[1], sys.argv[2]
    name, ins = sass(lib, kernel)
    print(name, len(ins), "instructions")
    addr2i = {a: i for i, (a, _, _, _) in enumerate(ins)}
    stall = [((hi >> 41) & 0xf) if hi is not None else 0 for _, _, _, hi in ins]
    wait = [((hi >> 52) & 0x3f) if hi is not None else 0 for _, _, _, hi in ins]
    if "--dump" in sys.argv:
        lo, hi_ = int(sys.argv[sys.argv.index("--dump") + 1]), int(sys.argv[sys.argv.index("--dump") + 2])
        for i in range(lo, hi_):
            print(i, f"st={stall[i]:2d} w={wait[i]:02x}", ins[i][1])
        return
    loops = []
    for i, (a, txt, _, _) in enumerate(ins):
        m = re.search(r"BRA(?:\.\w+)*\s+(?:\S+,\s*)?0x([0-9a-f]+)", txt)
        if m:
            t = int(m.group(1), 16)
            if t in addr2i and addr2i[t] <= i:
                loops.append((addr2i[t], i))
    loops.sort(key=lambda l: l[1] - l[0])
    for lo, hi_ in loops:
        n = hi_ - lo + 1
        cyc = sum(max(1, s) for s in stall[lo:hi_ + 1])
        ops = collections.Counter()
        for _, txt, _, _ in ins[lo:hi_ + 1]:
            t = txt.split()
            o = t[1] if t[0].startswith("@") else t[0]
            ops[o.split(".")[0]] += 1
        fp64 = sum(v for k, v in ops.items() if k in ("DFMA", "DADD", "DMUL", "DSETP"))
        nsb = sum(1 for w in wait[lo:hi_ + 1] if w)
        print(f"loop [{lo:5d},{hi_:5d}] n={n:4d} stall-sum={cyc:5d} cyc ({cyc / n:.2f}/instr) fp64={fp64} "
              f"(pipe floor {2 * fp64}) sb-waits={nsb}  " + " ".join(f"{k}:{v}" for k, v in ops.most_common(8)))


if __name__ == "__main__":
    main()
