cd $GRAFT_REPO_ROOT
timeout 300 python tools/chase_check.py 8 > gpurun_out/chase_check.json
python - <<'PY'
import json
d=json.load(open('gpurun_out/chase_check.json'))
print(d['rays'], d['ms'])
for k in ('wavefront_vs_chaser','wavefront_vs_mix1','mix1_vs_mix2'):
    print(k, d[k]['pixels'], [(f[0],f[1]) for f in d[k]['first']])
PY
timeout 300 python -m pytest tests -m gpu -x -q 2>&1 | tail -5
timeout 120 env CRT_DUMP_LANES=1 python tools/render_once.py --steps 2 2> gpurun_out/lanes_chase.txt
