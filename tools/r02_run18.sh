#!/bin/bash
# round 2, 2 GPUs: the bench line must complete (short timeouts: a hang must not burn the budget)
cd "${GRAFT_REPO_ROOT:-.}"; mkdir -p gpurun_out; O=gpurun_out
timeout 240 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29518 bench.py --gpus 2 --steps 20 --warmup 5 > $O/bench_n2c.json 2> $O/bench_n2c.err; echo "bench n2 rc=$?"; python - <<'PY'
import json
j=json.loads(open('gpurun_out/bench_n2c.json').readline())
print('ms/step', j['ms_per_step'], 'value', j['value'], 'check', j.get('exchange_check',{}).get('ok'))
t=j['other_configs']['cfg4-train']
print({k:t.get(k) for k in ('ms_per_step','cuda_graph_ms_per_step','cuda_graph_error','breakdown_ms','loss','folded_draw_kernels')})
for k,v in j.get('other_configs',{}).items(): print(k, {kk:vv for kk,vv in v.items() if kk in ('ms_per_step','roofline_frac','error')})
PY
grep -v -i 'warn' $O/bench_n2c.err | tail -n 5
