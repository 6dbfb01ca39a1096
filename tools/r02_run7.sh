#!/bin/bash
# 1 GPU: local cost of the peer-exchange protocol (world of one), blocking vs split-phase vs the plain entry point
cd "${GRAFT_REPO_ROOT:-.}"; mkdir -p gpurun_out; O=gpurun_out
run() { python bench.py --steps $1 --warmup 10 --no-cpu-baseline --no-other-configs --no-e2e $2 2>>$O/peer1.err | python -c "
import json,sys
j=json.loads(sys.stdin.readline()); print('$1 steps'.ljust(10), '$2'.ljust(30), 'ms/step %.4f check %s'%(j['ms_per_step'], (j.get('exchange_check') or {}).get('ok')))" | tee -a $O/peer1.txt; }
for k in 20 200; do
  run $k ""; run $k "--force-peer"; run $k "--force-peer --peer-blocking"; run $k ""; run $k "--force-peer"; run $k "--force-peer --peer-blocking"
done
for k in 20 200; do run $k "--config cfg4"; run $k "--config cfg4 --force-peer"; run $k "--config cfg4 --force-peer --peer-blocking"; done
