cd $GRAFT_REPO_ROOT
timeout 600 python -m pytest tests -m gpu -x -q 2>&1 | tail -4
bash tools/profile_r01.sh > gpurun_out/profile_r01.log 2>&1
tail -3 gpurun_out/profile_r01.log
