#!/bin/bash
# GPU-box recipe behind profiles/r02/*: every command runs plain first (must exit 0), then under ncu (B200_PROFILING.md).
cd $GRAFT_REPO_ROOT
set -x
python bench.py --steps 3 --warmup 3 --c4-throughput > gpurun_out/bench_c3.json 2> gpurun_out/bench_c3.err || exit 1
python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/bench_c3_reference.json 2>/dev/null
python bench.py --workload rtiow --spp 10 --steps 5 --warmup 3 > gpurun_out/bench_c1_rtiow10.json 2>/dev/null
python bench.py --workload rtiow --spp 100 --steps 5 --warmup 3 > gpurun_out/bench_c2_rtiow100.json 2>/dev/null
python bench.py --workload rtiow --spp 100 --impl reference --steps 2 --warmup 1 > gpurun_out/bench_c2_reference_derived.json 2>/dev/null
python bench.py --workload raybatch --steps 5 --warmup 3 > gpurun_out/bench_c5_raybatch.json 2>/dev/null
python bench.py --steps 3 --warmup 2 --slots 4 --no-c4 > gpurun_out/bench_c3_throughput_slots4.json 2>/dev/null
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_r02.csv python bench.py --steps 2 --warmup 3 --no-c4 > gpurun_out/ncu_launches.log 2>&1
python tools/render_once.py --ns 8 --steps 1 > gpurun_out/r_plain.log 2>&1 || exit 1
ncu --set full --import-source on --clock-control none -k regex:wideTraceKernel --launch-skip 20 --launch-count 1 -f -o gpurun_out/prof_wtrace_r02 python tools/render_once.py --ns 8 --steps 1 > gpurun_out/ncu_1.log 2>&1
ncu --set full --import-source on --clock-control none -k regex:meshShadeKernel --launch-skip 20 --launch-count 1 -f -o gpurun_out/prof_shade_r02 python tools/render_once.py --ns 8 --steps 1 > gpurun_out/ncu_2.log 2>&1
ncu --set full --import-source on --clock-control none -k regex:chaseKernel --launch-skip 3 --launch-count 1 -f -o gpurun_out/prof_chase_r02 python tools/render_once.py --nx 600 --ny 400 --ns 32 --steps 1 > gpurun_out/ncu_3.log 2>&1
python tools/prof_batch.py 1.0 22 0 3 > gpurun_out/prof_plain.log 2>&1 || exit 1
ncu --set full --import-source on --clock-control none -k regex:wideIntersectBatch -s 1 -c 1 -f -o gpurun_out/prof_widebatch_r02 python tools/prof_batch.py 1.0 22 0 3 > gpurun_out/ncu_4.log 2>&1
python tools/render_spheres_once.py > gpurun_out/rs_plain.log 2>&1 || exit 1
ncu --set full --import-source on --clock-control none -k regex:spheresMegaKernel --launch-skip 1 --launch-count 1 -f -o gpurun_out/prof_sph_mega_r02 python tools/render_spheres_once.py 100 1 > gpurun_out/ncu_5.log 2>&1
CRT_SPHERES_WAVEFRONT=1 ncu --set full --import-source on --clock-control none -k regex:extendSpheresBvhKernel --launch-skip 20 --launch-count 1 -f -o gpurun_out/prof_sph_extend_r02 python tools/render_spheres_once.py > gpurun_out/ncu_6.log 2>&1
CRT_SPHERES_WAVEFRONT=1 ncu --set full --import-source on --clock-control none -k regex:shadeSpheresKernel --launch-skip 20 --launch-count 1 -f -o gpurun_out/prof_sph_shade_r02 python tools/render_spheres_once.py > gpurun_out/ncu_7.log 2>&1
tail -2 gpurun_out/ncu_7.log
