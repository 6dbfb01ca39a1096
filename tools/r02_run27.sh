#!/bin/bash
# round 2, 4 GPUs, short timeout: the bench line after all round-2 work (scaling data point, hang check)
cd "${GRAFT_REPO_ROOT:-.}"; mkdir -p gpurun_out; O=gpurun_out
timeout 100 python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29521 bench.py --gpus 4 --steps 20 --warmup 5 > $O/bench_n4.json 2> $O/bench_n4.err; echo "bench n4 rc=$?"; python - <<'PY'
import json
txt=open('gpurun_out/bench_n4.json').read()
j=json.loads([l for l in txt.splitlines() if l.startswith('{')][-1])
print('ms/step', j['ms_per_step'], 'value', j['value'], 'check', j.get('exchange_check',{}).get('ok'))
t=j['other_configs']['cfg4-train']
print({k:t.get(k) for k in ('ms_per_step','breakdown_ms','loss')})
print({k:v.get('ms_per_step') for k,v in j['other_configs'].items()})
PY
