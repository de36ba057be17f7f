cd $GRAFT_REPO_ROOT
timeout 120 env CRT_DUMP_LANES=1 python tools/render_once.py --steps 1 2> gpurun_out/lanes_chase.txt
timeout 300 python -m pytest tests -m gpu -x -q 2>&1 | tail -3
run() { echo "=== $*"; env "$@" python tools/render_once.py --steps 2 | tail -1 | cut -c1-60; }
run CRT_X=1
run CRT_CHASE_LAG_PCT=30
run CRT_CHASE_LAG_MAX_PCT=95
run CRT_CHASE_MOVE_ALL=65536
run CRT_CHASE_MOVE_ALL=8192
run CRT_CHASE_CAPACITY=8192
run CRT_CHASE_CAPACITY=32768
run CRT_CHASE_LAST_WAVE=4736
run CRT_LANE_A_SHADE_BLOCKS=444
