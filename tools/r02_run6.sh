#!/bin/bash
cd "${GRAFT_REPO_ROOT:-.}"; mkdir -p gpurun_out; O=gpurun_out
N=${1:-2}
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511"
timeout 600 python -m pytest tests/test_peer_gpu.py tests/test_dp_fit_gpu.py -m gpu -x -q > $O/pytest_peer_n$N.log 2>&1; echo "pytest rc=$?"; tail -3 $O/pytest_peer_n$N.log
timeout 600 $TR bench.py --gpus $N --steps 20 --warmup 5 > $O/bench_n$N.json 2> $O/bench_n$N.err; echo "bench n$N rc=$?"
timeout 600 $TR bench.py --gpus $N --steps 20 --warmup 5 --peer-blocking --no-other-configs > $O/bench_n${N}_blocking.json 2> $O/bench_n${N}_blocking.err; echo "rc=$?"
timeout 600 $TR bench.py --gpus $N --steps 20 --warmup 5 --exchange nccl --no-other-configs > $O/bench_n${N}_nccl.json 2> $O/bench_n${N}_nccl.err; echo "rc=$?"
timeout 600 $TR bench.py --gpus $N --steps 200 --warmup 20 --no-other-configs > $O/bench_n${N}_200.json 2> /dev/null; echo "rc=$?"
python - <<PY
import json
for f in ('bench_n$N','bench_n${N}_blocking','bench_n${N}_nccl','bench_n${N}_200'):
    try:
        j=json.loads([l for l in open('$O/%s.json'%f) if l.startswith('{')][0])
        print(f, 'ms/step %.4f'%j['ms_per_step'], 'kernel_ms %.4f'%j['roofline']['kernel_ms'], 'check', j['exchange_check']['ok'], 'e2e %.3e'%j['e2e']['value'])
        for k,v in (j.get('other_configs') or {}).items():
            print('   ',k, {kk:v[kk] for kk in ('ms_per_step','roofline_frac','breakdown_ms','error') if kk in v})
    except Exception as e: print(f, 'FAILED', e)
PY
