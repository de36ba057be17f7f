cd $GRAFT_REPO_ROOT
run() { echo "=== $*"; env "$@" python tools/render_once.py --steps 2 | tail -1 | cut -c1-60; }
run CRT_B200_LIB=build/libcrt_b200_head.so
run CRT_X=1
run CRT_B200_LIB=build/libcrt_b200_head.so
run CRT_X=1
run CRT_B200_LIB=build/libcrt_b200_head.so CRT_EXPRESS_LANE=0
run CRT_EXPRESS_LANE=0
echo "=== head batch"; CRT_B200_LIB=build/libcrt_b200_head.so python tools/batch_halves.py
echo "=== new batch"; python tools/batch_halves.py
timeout 600 python -m pytest tests -m gpu -x -q 2>&1 | tail -4
