#!/bin/bash
# round 2, 2 GPUs: data-parallel tests + the bench line (captured Bayesian step with its all-reduce)
cd "${GRAFT_REPO_ROOT:-.}"; mkdir -p gpurun_out; O=gpurun_out
timeout 900 python -m pytest tests/test_dp_fit_gpu.py tests/test_peer_gpu.py -m gpu -x -q > $O/pytest_n2b.log 2>&1; echo "pytest n2 rc=$?"; tail -n 5 $O/pytest_n2b.log
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29517 bench.py --gpus 2 --steps 20 --warmup 5 > $O/bench_n2b.json 2> $O/bench_n2b.err; echo "bench n2 rc=$?"; python - <<'PY'
import json
j=json.loads(open('gpurun_out/bench_n2b.json').readline())
print('ms/step', j['ms_per_step'], 'value', j['value'], 'check', j.get('exchange_check',{}).get('ok'))
t=j['other_configs']['cfg4-train']
print({k:t.get(k) for k in ('ms_per_step','cuda_graph_ms_per_step','cuda_graph_error','breakdown_ms','loss','folded_draw_kernels')})
for k,v in j.get('other_configs',{}).items(): print(k, {kk:vv for kk,vv in v.items() if kk in ('ms_per_step','roofline_frac','error')})
PY
tail -n 3 $O/bench_n2b.err
