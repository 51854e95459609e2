#!/bin/bash
# tools/capture_profiles.sh — every capture behind profiles/r02_*.md in one GPU call (run under gpurun from the repo root):
#   ncu --set full of the solve kernel on (a) bench.py's default step, (b) the saturated launch, (c) one problem alone;
#   the launch list of a bench run; compute-sanitizer x 4 on tools/sanitize.py.  Then: python tools/make_profiles.py (CPU).
set -x
O=gpurun_out
NCU="ncu --set full --clock-control none --import-source on -k regex:nmpc_solve --launch-count 1 -f"
$NCU --launch-skip 3 -o $O/r02_full_config2 python bench.py --steps 2 --warmup 3 --extra off > $O/r02_ncu_config2.log 2>&1
$NCU --launch-skip 1 -o $O/r02_saturated python tools/sat.py > $O/r02_saturated.log 2>&1
$NCU --launch-skip 1 -o $O/r02_lone python tools/lone.py 1207 2 > $O/r02_lone.log 2>&1
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $O/r02_launches.csv python bench.py --steps 2 --warmup 3 --extra off > $O/r02_launches.log 2>&1
for t in memcheck synccheck initcheck racecheck; do
  timeout 600 compute-sanitizer --tool $t --print-limit 400 python tools/sanitize.py > $O/r02_san_$t.log 2>&1
done
tail -2 $O/r02_san_*.log
