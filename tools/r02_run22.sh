#!/bin/bash
# round 2, 2 GPUs, short timeouts: data-parallel tests and the bench line after the weight kernels / fused Adam
cd "${GRAFT_REPO_ROOT:-.}"; mkdir -p gpurun_out; O=gpurun_out
timeout 200 python -m pytest tests/test_dp_fit_gpu.py tests/test_peer_gpu.py -m gpu -x -q > $O/pytest_n2c.log 2>&1; echo "pytest n2 rc=$?"; tail -n 4 $O/pytest_n2c.log | cut -c1-200
timeout 150 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29519 bench.py --gpus 2 --steps 20 --warmup 5 > $O/bench_n2d.json 2> $O/bench_n2d.err; echo "bench n2 rc=$?"; python - <<'PY'
import json
txt=open('gpurun_out/bench_n2d.json').read()
j=json.loads([l for l in txt.splitlines() if l.startswith('{')][-1])
print('ms/step', j['ms_per_step'], 'value', j['value'], 'check', j.get('exchange_check',{}).get('ok'))
t=j['other_configs']['cfg4-train']
print({k:t.get(k) for k in ('ms_per_step','cuda_graph_ms_per_step','breakdown_ms','loss','folded_draw_kernels')})
print({k:v.get('ms_per_step') for k,v in j['other_configs'].items()})
PY
grep -v -i 'warn\|\*\*\*\|OMP_NUM' $O/bench_n2d.err | tail -n 5
