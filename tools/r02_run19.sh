#!/bin/bash
# round 2: variational-weights kernel + fused Adam in the Bayesian step; whole GPU suite, step profile, bench line
cd "${GRAFT_REPO_ROOT:-.}"; mkdir -p gpurun_out; O=gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q > $O/pytest_gpu8.log 2>&1; echo "pytest all rc=$?"; tail -n 6 $O/pytest_gpu8.log
timeout 300 python tools/bayes_step_profile.py 2>&1 | grep -v Warn > $O/bayes_step_profile3.txt; head -n 14 $O/bayes_step_profile3.txt
timeout 600 python bench.py --steps 20 --warmup 5 > $O/bench_r02d.json 2> $O/bench_r02d.err; echo "bench rc=$?"; python - <<'PY'
import json
txt=open('gpurun_out/bench_r02d.json').read()
j=json.loads([l for l in txt.splitlines() if l.startswith('{')][-1])
print('ms/step', j['ms_per_step'], 'frac', j['roofline']['frac'])
t=j['other_configs']['cfg4-train']
print({k:t.get(k) for k in ('ms_per_step','cuda_graph_ms_per_step','cuda_graph_error','loss')})
print({k:v.get('value', v.get('ms_per_step')) for k,v in j['other_configs'].items()})
PY
