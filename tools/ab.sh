cd $GRAFT_REPO_ROOT
timeout 300 python tools/chase_check.py 8 > gpurun_out/chase_check.json
python - <<'PY'
import json
d=json.load(open('gpurun_out/chase_check.json'))
print(d['rays'], d['ms'])
for k in ('wavefront_vs_chaser','wavefront_vs_mix1','mix1_vs_mix2'):
    print(k, d[k]['pixels'], [(f[0],f[1]) for f in d[k]['first']])
PY
timeout 600 python -m pytest tests -m gpu -x -q 2>&1 | tail -4
run() { echo "=== $*"; env "$@" python tools/render_once.py --steps 2 | tail -1 | cut -c1-60; }
run CRT_X=1
run CRT_CHASE_WAVE=296
run CRT_CHASE_WAVE=148
run CRT_CHASE_WAVE=296 CRT_CHASE_TARGET=6000
run CRT_CHASE_TARGET=6000
run CRT_CHASE_TARGET=8000 CRT_CHASE_CAPACITY=32768
run CRT_CHASE_EXCLUSIVE_PCT=12
run CRT_CHASE_EXCLUSIVE_SLOTS=2 CRT_CHASE_EXCLUSIVE_PCT=12
timeout 120 env CRT_DUMP_LANES=1 python tools/render_once.py --steps 1 2> gpurun_out/lanes_chase2.txt
