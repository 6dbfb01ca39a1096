#!/bin/bash
# round 2: Bayesian step -- no validation syncs, CUDA-graph replay; tests + profile + bench line
cd "${GRAFT_REPO_ROOT:-.}"; mkdir -p gpurun_out; O=gpurun_out
timeout 900 python -m pytest tests/test_draws_gpu.py tests/test_estimators_gpu.py -m gpu -x -q > $O/pytest_draws.log 2>&1; echo "pytest rc=$?"; tail -n 12 $O/pytest_draws.log
timeout 300 python tools/bayes_step_profile.py 2>&1 | grep -v Warn > $O/bayes_step_profile2.txt; head -n 12 $O/bayes_step_profile2.txt
timeout 900 python bench.py --steps 20 --warmup 5 > $O/bench_r02c.json 2> $O/bench_r02c.err; echo "bench rc=$?"; python - <<'PY'
import json
j=json.loads(open('gpurun_out/bench_r02c.json').readline())
print('ms/step', j['ms_per_step'], 'frac', j['roofline']['frac'])
print(json.dumps(j['other_configs']['cfg4-train'], indent=1))
PY
tail -n 3 $O/bench_r02c.err
