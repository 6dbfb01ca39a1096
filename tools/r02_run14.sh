#!/bin/bash
# round 2: whole GPU suite + default bench line after the fused-kernel work
cd "${GRAFT_REPO_ROOT:-.}"; mkdir -p gpurun_out; O=gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q > $O/pytest_gpu6.log 2>&1; echo "pytest all rc=$?"; tail -n 6 $O/pytest_gpu6.log
timeout 900 python bench.py --steps 20 --warmup 5 > $O/bench_r02b.json 2> $O/bench_r02b.err; echo "bench rc=$?"; python - <<'PY'
import json
j=json.loads(open('gpurun_out/bench_r02b.json').readline())
print('ms/step', j['ms_per_step'], 'frac', j['roofline']['frac'], 'e2e', j['e2e']['value'])
print('fused_dense_step', j.get('fused_dense_step'))
print('fused_dense_mdn_step', j.get('fused_dense_mdn_step'))
for k,v in j.get('other_configs',{}).items(): print(k, {kk:vv for kk,vv in v.items() if kk in ('ms_per_step','roofline_frac','error','value')})
PY
tail -n 3 $O/bench_r02b.err
