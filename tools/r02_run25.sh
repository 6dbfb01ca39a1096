#!/bin/bash
# round 2: the Bayesian step's network part as one library call -- tests, profile, bench line
cd "${GRAFT_REPO_ROOT:-.}"; mkdir -p gpurun_out; O=gpurun_out
timeout 900 python -m pytest tests/test_draws_gpu.py tests/test_estimators_gpu.py tests/test_dense_gpu.py -m gpu -q > $O/pytest_onecall.log 2>&1; echo "pytest rc=$?"; tail -n 10 $O/pytest_onecall.log | cut -c1-220
timeout 300 python tools/bayes_step_profile.py 2>&1 | grep -v Warn > $O/bayes_step_profile5.txt; head -n 10 $O/bayes_step_profile5.txt | cut -c1-150
timeout 600 python bench.py --steps 20 --warmup 5 > $O/bench_r02f.json 2> $O/bench_r02f.err; echo "bench rc=$?"; python - <<'PY'
import json
txt=open('gpurun_out/bench_r02f.json').read()
j=json.loads([l for l in txt.splitlines() if l.startswith('{')][-1])
print('ms/step', j['ms_per_step'], 'frac', j['roofline']['frac'])
t=j['other_configs']['cfg4-train']
print({k:t.get(k) for k in ('ms_per_step','cuda_graph_ms_per_step','cuda_graph_error','loss')})
PY
