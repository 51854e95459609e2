#!/usr/bin/env python
"""tools/make_profiles.py — compose the committed round-2 profile summaries from what a GPU run left in gpurun_out/
(r02_full_config2 / r02_saturated / r02_lone .ncu-rep, r02_launches.csv, r02_san_*.log, r02_sweep.json, r02_fleet.json):
    python tools/make_profiles.py
writes profiles/r02_ncu_summary.md, profiles/r02_sanitizers.md, profiles/traffic.json, profiles/r02_launches_config2.csv,
profiles/r02_sweep.json, profiles/r02_fleet.json."""
import collections
import csv
import json
import os
import re
import shutil
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
G = os.path.join(ROOT, "gpurun_out")
P = os.path.join(ROOT, "profiles")


def summary(rep):
    return subprocess.run([sys.executable, os.path.join(ROOT, "tools", "ncu_summary.py"), rep, "--json", rep + ".json"],
                          capture_output=True, text=True).stdout


def launches():
    rows = [r for r in csv.reader(open(os.path.join(G, "r02_launches.csv"))) if len(r) > 5]
    ix = {h: i for i, h in enumerate(rows[0])}
    agg = collections.defaultdict(lambda: [0, 0.0])
    for r in rows[1:]:
        if r[ix["Metric Name"]] != "gpu__time_duration.sum":
            continue
        v, u = float(r[ix["Metric Value"]]), r[ix["Metric Unit"]]
        v = v / 1e6 if u == "ns" else (v / 1e3 if u == "us" else v)
        k = re.sub(r"\(.*", "", r[ix["Kernel Name"]])[:70]
        agg[k][0] += 1
        agg[k][1] += v
    tot = sum(v[1] for v in agg.values())
    out = "| kernel | launches | total under ncu | share |\n|---|---|---|---|\n"
    for k, (n, t) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
        out += f"| `{k}` | {n} | {t:.3f} ms | {100 * t / tot:.2f} % |\n"
    return out


def main():
    import bench
    md = ["# Round 2 — ncu evidence for `nmpc_solve_kernel<8,3>` (B200, sm_100a)\n",
          "Captures: `ncu --set full --clock-control none --import-source on -k regex:nmpc_solve` (one launch after the warm-up\n"
          "launches) of (a) `bench.py` default = BASELINE configs[1] (B=4096, map 3, N=20, cold start), (b) a saturated\n"
          "synthetic batch (B=32 768), (c) the hardest problem of config 2 alone (`tools/lone.py 1207`); launch list of (a) with\n"
          "`--metrics gpu__time_duration.sum`.  Tables made by `tools/ncu_summary.py` / `tools/make_profiles.py` from the\n"
          "`.ncu-rep` files.  Numbers taken under the profiler are not bench values.\n",
          f"Kernel source hash (bench.py `roofline.kernel_source_hash`): `{bench.kernel_source_hash()}`\n",
          "## Launch list of one `bench.py --steps 2 --warmup 3` run (share of a step)\n", launches(),
          "\nEvery bench step = probe (one evaluation per problem) + scan + scatter (longest-first order) + one\n"
          "`nmpc_solve_kernel` (grid 148 x 384 threads); `gpu_launches` = 4 per step.\n"]
    for title, rep in (("(a) BASELINE config 2, B = 4096", "r02_full_config2"), ("(b) saturated, synthetic B = 32 768", "r02_saturated"),
                       ("(c) one problem alone (4 7xx inner iterations)", "r02_lone")):
        path = os.path.join(G, rep + ".ncu-rep")
        if os.path.exists(path):
            md += [f"\n## {title}\n", summary(path)]
    with open(os.path.join(P, "r02_ncu_summary.md"), "w") as f:
        f.write("\n".join(md))
    # traffic.json keyed to the kernel sources
    j = json.load(open(os.path.join(G, "r02_full_config2.ncu-rep.json")))
    rd = float(j["metrics"]["dram__bytes_read.sum"]["value"]) * {"Mbyte": 1e6, "Kbyte": 1e3, "byte": 1, "Gbyte": 1e9}[j["metrics"]["dram__bytes_read.sum"]["unit"]]
    wr = float(j["metrics"]["dram__bytes_write.sum"]["value"]) * {"Mbyte": 1e6, "Kbyte": 1e3, "byte": 1, "Gbyte": 1e9}[j["metrics"]["dram__bytes_write.sum"]["unit"]]
    json.dump({"source": "ncu --set full --clock-control none --import-source on -k regex:nmpc_solve --launch-skip 3 --launch-count 1, "
                         "python bench.py --steps 2 --warmup 3 (gpurun_out/r02_full_config2.ncu-rep, nmpc_solve_kernel<8,3>)",
               "kernel_source_hash": bench.kernel_source_hash(),
               "captures": [{"workload": "config2", "batch": 4096, "dram_bytes_read": int(rd), "dram_bytes_write": int(wr),
                             "dram_bytes_per_launch": int(rd + wr), "algorithmic_bytes_per_launch": 4128 * 4096,
                             "note": "reads: P rows + U0/Y0; the results were still in L2 when the kernel ended"}]},
              open(os.path.join(P, "traffic.json"), "w"), indent=1)
    shutil.copy(os.path.join(G, "r02_launches.csv"), os.path.join(P, "r02_launches_config2.csv"))
    for n in ("r02_sweep.json", "r02_fleet.json"):
        if os.path.exists(os.path.join(G, n)):
            shutil.copy(os.path.join(G, n), os.path.join(P, n))
    # sanitizers
    san = ["# Round 2 — compute-sanitizer on `tools/sanitize.py` (small batches N=20 / N=40 with obstacles in the way, fleet steps "
           "with the device sampler)\n", "| tool | summary |\n|---|---|"]
    for t in ("memcheck", "racecheck", "synccheck", "initcheck"):
        p = os.path.join(G, f"r02_san_{t}.log")
        if os.path.exists(p):
            txt = open(p).read()
            m = re.findall(r"(ERROR SUMMARY: .*|RACECHECK SUMMARY: .*)", txt)
            san.append(f"| {t} | {m[-1] if m else 'no summary line'} |")
            if t == "racecheck":
                kinds = collections.Counter(re.sub(r"\+0x[0-9a-f]+", "", l.split("between ")[1].strip())
                                            for l in txt.splitlines() if "Race reported between" in l)
                rc = kinds
    # the same tool with the helper warps switched off (NMPC_B200_HELP_MAX_WAVES=0): the solver itself
    p0 = os.path.join(G, "r02_san_racecheck_nohelp.log")
    if os.path.exists(p0):
        m = re.findall(r"(RACECHECK SUMMARY: .*)", open(p0).read())
        san.append(f"| racecheck, helper warps off (`NMPC_B200_HELP_MAX_WAVES=0`) | {m[-1] if m else 'no summary line'} |")
    san.append("\nracecheck reports by first access:\n")
    san += [f"* {n} x `{k}`" for k, n in rc.most_common()]
    san.append("""
**Warnings** (as before the helper warps existed): scalar slots of the arena header (`sget` / `sput` / `iget` / `iput` in
`csrc/nmpc_device.cuh`: the warp-uniform solver state).  All 32 lanes of the owning warp execute the same store with the
same value and every lane later reads the value back; racecheck sees lane A's store and lane B's load of one address
without a barrier in between.  By construction the value a lane reads is the one it stored itself (program order) or an
identical one.  The cross-group vector exchanges inside a warp (`Warp::st` / `Warp::ld`) are separated by `__syncwarp()`
and produce no report.

**Errors**: every one is an access pair between an owner warp and its helper warp (DESIGN.md section 5, "Helper warps") —
with the helper warps off the same workload reports none.  The two warps synchronise through the mailbox words with
`st.release.cta` / `ld.acquire.cta` (and `__syncwarp()` inside each warp before the release and after the acquire);
racecheck only models barriers, so it reports every pair of accesses the flags order:

| first access | second access | what it is | what orders it |
|---|---|---|---|
| read `lds2` (helper: `form_trials` loads the owner's V_U, V_FPR, V_DIR) | write `sts2_if` (owner: `Warp::st` of those vectors in PH_STEP_BEGIN / at an accepted trial) | the helper reads the trial inputs of post s | the owner stores them, `__syncwarp()`, then `st.release` of `seq = s`; the helper's `ld.acquire` of `seq` precedes its loads.  The owner overwrites V_U only after it has either taken the helper's results for s or accepted a trial of its own call — in the second case a late helper may read a half-written V_U, and its results for s are never taken (`done == seq` is only waited for when the owner's call accepted nothing) |
| write `sts2` / `sts1` (helper: PH_HELP stores x, grad, gradient step, half step per trial and (psi, lhs) into its OWN arena) | read `lds2` / `lds1` (owner: `take_from_helper`) | the owner takes the results of post s | helper: stores, `__syncwarp()`, `st.release` of `done = s`; owner: `ld.acquire` until `done == s`, then reads.  The helper overwrites its arena only after it has seen `seq != s`, which the owner writes after it has finished copying |
| write `st_release` | read `ld_acquire` / `ldsi` | the mailbox words themselves (`seq`, `done`, `state`, `helper`) | these ARE the synchronisation |
| read `ldsi` | write `stsi` / atomic | a retiring warp reads the iteration counters of the solving warps of its CTA to choose the one it helps (any value is acceptable: a heuristic), and the owner reads its `helper` word, which the helper sets once with a compare-and-swap | nothing needs to |
| read `lds1` (helper: the owner's step size, penalty and problem data inside the evaluation) | write `sts1` (owner: `sput`) | the helper evaluates with the owner's header | those slots change only in `lip_halve` and at the end of an inner solve, and the helper's results of a post are not taken after either |

memcheck, synccheck (every `__syncwarp` / `__shfl_sync` / vote with the full mask, also in helper mode) and initcheck are clean
with the helper warps on.""")
    with open(os.path.join(P, "r02_sanitizers.md"), "w") as f:
        f.write("\n".join(san) + "\n")
    print("profiles written")


if __name__ == "__main__":
    main()
