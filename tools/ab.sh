cd $GRAFT_REPO_ROOT
timeout 600 python -m pytest tests -m gpu -x -q 2>&1 | tail -4
python tools/parity_report.py --detail 0.25 --tex 64 --nx 1200 --ny 800 --ns 100 --rays 65536 --steps 2 --spheres 2>&1 | grep SPHERES | cut -c1-600
CRT_SPHERES_BRUTE=1 python bench.py --workload rtiow --steps 2 --warmup 3 2>/dev/null | cut -c1-700
python bench.py --workload rtiow --steps 2 --warmup 3 2>/dev/null | cut -c1-900
python bench.py --workload rtiow --impl reference --steps 2 --warmup 1 2>/dev/null | cut -c1-500
