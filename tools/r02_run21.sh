#!/bin/bash
# round 2: whole GPU suite + smoke + default bench
cd "${GRAFT_REPO_ROOT:-.}"; mkdir -p gpurun_out; O=gpurun_out
timeout 1500 python -m pytest tests -m gpu -q > $O/pytest_gpu10.log 2>&1; echo "pytest all rc=$?"; tail -n 8 $O/pytest_gpu10.log | cut -c1-200
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > $O/smoke2.log 2>&1; echo "smoke rc=$?"; tail -n 2 $O/smoke2.log
timeout 600 python bench.py --steps 20 --warmup 5 > $O/bench_r02e.json 2> $O/bench_r02e.err; echo "bench rc=$?"; python - <<'PY'
import json
txt=open('gpurun_out/bench_r02e.json').read()
j=json.loads([l for l in txt.splitlines() if l.startswith('{')][-1])
print('ms/step', j['ms_per_step'], 'frac', j['roofline']['frac'], 'e2e', j['e2e']['value'])
t=j['other_configs']['cfg4-train']
print({k:t.get(k) for k in ('ms_per_step','cuda_graph_ms_per_step','cuda_graph_error','loss')})
print({k:v.get('value', v.get('ms_per_step')) for k,v in j['other_configs'].items()})
print(j['fused_dense_step']['us_per_step'], j['fused_dense_mdn_step'].get('us_per_step'), j['fused_dense_mdn_step'].get('speedup_vs_unfused'))
c=j['other_configs']['cfg1-pipeline']; print({k:c[k] for k in c if k.endswith('_us')})
PY
