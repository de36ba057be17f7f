#!/bin/bash
# GPU-box recipe behind profiles/r01/*: plain bench first (must exit 0), then the ncu launch list of the same command,
# then one --set full capture per hot kernel (B200_PROFILING.md). Outputs land in gpurun_out/.
cd $GRAFT_REPO_ROOT
set -x
python bench.py --steps 2 --warmup 3 > gpurun_out/bench_plain.json 2> gpurun_out/bench_plain.err || exit 1
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_r01.csv python bench.py --steps 2 --warmup 3 > gpurun_out/ncu_launches.log 2>&1
ncu --set full --import-source on --clock-control none -k regex:traceKernel --launch-skip 20 --launch-count 1 -f -o gpurun_out/prof_trace_r01 python tools/render_once.py --ns 8 --steps 1 > gpurun_out/ncu_trace.log 2>&1
ncu --set full --import-source on --clock-control none -k regex:meshShadeKernel --launch-skip 20 --launch-count 1 -f -o gpurun_out/prof_shade_r01 python tools/render_once.py --ns 8 --steps 1 > gpurun_out/ncu_shade.log 2>&1
ncu --set full --import-source on --clock-control none -k regex:chaseKernel --launch-skip 3 --launch-count 1 -f -o gpurun_out/prof_chase_r01 python tools/render_once.py --nx 600 --ny 400 --ns 32 --steps 1 > gpurun_out/ncu_chase.log 2>&1
tail -3 gpurun_out/ncu_chase.log
