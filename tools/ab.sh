cd $GRAFT_REPO_ROOT
python -m pytest tests -m gpu -x -q 2>&1 | tail -3
python tools/e2e_breakdown.py 100
CRT_TIMING=1 python tools/e2e_breakdown.py 8 2>&1 | tail -22
