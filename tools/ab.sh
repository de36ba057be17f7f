cd $GRAFT_REPO_ROOT
timeout 600 python -m pytest tests -m gpu -x -q 2>&1 | tail -8
python tools/render_once.py --steps 2 | cut -c1-100
