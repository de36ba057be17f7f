set -x
cd $GRAFT_REPO_ROOT
for d in 0 8 16 24 33; do
  echo "=== drain $d lanes"; CRT_TRACE_DRAIN=$d python tools/render_once.py --steps 2
done
for d in 0 16 33; do
  echo "=== drain $d nolanes"; CRT_EXPRESS_LANE=0 CRT_TRACE_DRAIN=$d python tools/render_once.py --steps 1
done
CRT_TRACE_DRAIN=33 python -m pytest tests -m gpu -x -q 2>&1 | tail -5
