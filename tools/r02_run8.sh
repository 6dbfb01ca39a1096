#!/bin/bash
# 1 GPU: peer tests (world of one) + local cost of the exchange protocols
cd "${GRAFT_REPO_ROOT:-.}"; mkdir -p gpurun_out; O=gpurun_out
timeout 300 python -m pytest tests/test_peer_gpu.py -m gpu -x -q > $O/pytest_peer_v4.log 2>&1; echo "pytest rc=$?"; tail -3 $O/pytest_peer_v4.log
run() { python bench.py --steps $1 --warmup 10 --no-cpu-baseline --no-other-configs --no-e2e $2 2>>$O/peer1.err | python -c "
import json,sys
j=json.loads(sys.stdin.readline()); print('$1 steps'.ljust(10), '$2'.ljust(30), 'ms/step %.4f check %s'%(j['ms_per_step'], (j.get('exchange_check') or {}).get('ok')))" | tee -a $O/peer1_v4.txt; }
for k in 20 200; do
  run $k ""; run $k "--force-peer"; run $k "--force-peer --peer-blocking"
done
for k in 20 200; do run $k "--config cfg4"; run $k "--config cfg4 --force-peer"; run $k "--config cfg4 --force-peer --peer-blocking"; done
