#!/usr/bin/env python
"""tools/ncu_summary.py — turn an `ncu --set full --import-source on` report of the solve kernel into the markdown /
json pieces committed under profiles/ (no GPU needed: reads the .ncu-rep with `ncu -i`).

    python tools/ncu_summary.py gpurun_out/x.ncu-rep [--json out.json]

Prints: headline metrics (duration, registers, DRAM bytes, issue / FP64-pipe utilisation, active cycles, thread
utilisation), the warp-stall breakdown, the executed-instruction opcode histogram of the SASS, and the hot
footprint (instructions that carry 99 % of the executed instructions)."""
import collections
import csv
import io
import json
import subprocess
import sys

WANT = {
    "gpu__time_duration.sum": "kernel duration",
    "launch__registers_per_thread": "registers per thread",
    "launch__grid_size": "grid",
    "launch__block_size": "block",
    "dram__bytes_read.sum": "DRAM bytes read",
    "dram__bytes_write.sum": "DRAM bytes written",
    "smsp__inst_executed.sum": "warp instructions executed",
    "smsp__thread_inst_executed_per_inst_executed.ratio": "active threads per instruction",
    "smsp__issue_active.avg.pct_of_peak_sustained_active": "issue slots busy (while active)",
    "sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active": "FP64 pipe (while active)",
    "smsp__cycles_active.avg": "SMSP cycles active (avg)",
    "sm__cycles_elapsed.max": "SM cycles elapsed",
    "sm__warps_active.avg.pct_of_peak_sustained_active": "warps active (of 64 per SM)",
    "sass__inst_executed_local_loads": "local (spill) loads",
    "sass__inst_executed_local_stores": "local (spill) stores",
}


def ncu(rep, *args):
    return subprocess.run(["ncu", "-i", rep, *args], capture_output=True, text=True).stdout


def main():
    rep = sys.argv[1]
    rows = list(csv.reader(io.StringIO(ncu(rep, "--page", "raw", "--csv"))))
    h, u, v = rows[0], rows[1], rows[2]
    raw = {k: (v[i], u[i]) for i, k in enumerate(h)}
    out = {"report": rep, "kernel": raw.get("Kernel Name", ("?", ""))[0], "metrics": {}, "stalls_per_issue": {}}
    print(f"kernel: {out['kernel']}")
    print("| metric | value |\n|---|---|")
    for k, name in WANT.items():
        if k in raw:
            out["metrics"][k] = {"value": raw[k][0], "unit": raw[k][1]}
            print(f"| {name} (`{k}`) | {raw[k][0]} {raw[k][1]} |")
    try:
        act = float(raw["smsp__cycles_active.avg"][0]) / float(raw["sm__cycles_elapsed.max"][0])
        out["metrics"]["active_over_elapsed"] = act
        print(f"| SMSP active / elapsed | {act:.3f} |")
    except (KeyError, ValueError):
        pass
    print("\n| stall reason | warps stalled per issued instruction |\n|---|---|")
    st = []
    for k in raw:
        if k.startswith("smsp__average_warps_issue_stalled_") and k.endswith("_per_issue_active.ratio"):
            try:
                st.append((float(raw[k][0]), k[len("smsp__average_warps_issue_stalled_"):-len("_per_issue_active.ratio")]))
            except ValueError:
                pass
    for val, name in sorted(st, reverse=True):
        if val >= 0.02:
            out["stalls_per_issue"][name] = val
            print(f"| {name} | {val:.3f} |")
    src = list(csv.reader(io.StringIO(ncu(rep, "--page", "source", "--csv"))))
    hdr = src[1]
    ix = {c: i for i, c in enumerate(hdr)}
    data = src[2:]
    ops, tot = collections.Counter(), 0
    ex = []
    for r in data:
        n = int(r[ix["Instructions Executed"]])
        t = r[ix["Source"]].split()
        o = (t[1] if t[0].startswith("@") else t[0]).split(".")[0]
        ops[o] += n
        tot += n
        ex.append(n)
    print(f"\nSASS: {len(data)} instructions in the kernel; executed-instruction histogram (share of {tot:.3e}):\n")
    print("| opcode | share |\n|---|---|")
    out["opcode_share"] = {}
    for o, n in ops.most_common(18):
        out["opcode_share"][o] = n / tot
        print(f"| {o} | {100 * n / tot:.2f} % |")
    fp64 = sum(ops[o] for o in ("DFMA", "DADD", "DMUL", "DSETP"))
    print(f"\nFP64-pipe instructions (DFMA+DADD+DMUL+DSETP): {100 * fp64 / tot:.1f} % of the executed instructions")
    srt = sorted(ex, reverse=True)
    acc, n99 = 0, 0
    for n in srt:
        acc += n
        n99 += 1
        if acc > 0.99 * tot:
            break
    out["hot_instructions_99pct"] = n99
    print(f"hot footprint: {n99} SASS instructions ({n99 * 16 / 1024:.1f} KB) carry 99 % of the executed instructions")
    if "--json" in sys.argv:
        with open(sys.argv[sys.argv.index("--json") + 1], "w") as f:
            json.dump(out, f, indent=1)


if __name__ == "__main__":
    main()
