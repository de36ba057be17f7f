cd $GRAFT_REPO_ROOT
for v in "" _top6 _top8 _top9; do
  lib=build/libcrt_b200$v.so
  echo "=== $lib batch"; CRT_B200_LIB=$lib python tools/batch_halves.py
  echo "=== $lib nochase"; CRT_EXPRESS_LANE=0 CRT_B200_LIB=$lib python tools/render_once.py --steps 1 | tail -1 | cut -c1-60
  echo "=== $lib frame"; CRT_B200_LIB=$lib python tools/render_once.py --steps 2 | tail -1 | cut -c1-60
done
CRT_B200_LIB=build/libcrt_b200_top8.so timeout 600 python -m pytest tests -m gpu -x -q 2>&1 | tail -3
