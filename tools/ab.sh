cd $GRAFT_REPO_ROOT
timeout 600 compute-sanitizer --tool memcheck --error-exitcode 7 python tools/render_once.py --nx 240 --ny 160 --ns 12 --detail 0.25 --tex 64 --steps 1 2>&1 | tail -8
echo "memcheck rc=$?"
timeout 600 compute-sanitizer --tool racecheck --error-exitcode 7 python tools/render_once.py --nx 160 --ny 100 --ns 8 --detail 0.25 --tex 64 --steps 1 2>&1 | tail -8
echo "racecheck rc=$?"
timeout 600 compute-sanitizer --tool initcheck --error-exitcode 7 python tools/render_once.py --nx 160 --ny 100 --ns 8 --detail 0.25 --tex 64 --steps 1 2>&1 | tail -12
echo "initcheck rc=$?"
