#!/bin/bash
cd "${GRAFT_REPO_ROOT:-.}"; mkdir -p gpurun_out; O=gpurun_out
run() { python bench.py --steps $1 --warmup 10 --no-cpu-baseline --no-other-configs --no-e2e $2 2>>$O/peer1.err | python -c "
import json,sys
j=json.loads(sys.stdin.readline()); print('$1 steps'.ljust(10), '$2'.ljust(30), 'ms/step %.4f flush_ms %s check %s'%(j['ms_per_step'], j.get('flush_ms'), (j.get('exchange_check') or {}).get('ok')))" | tee -a $O/peer1_v4b.txt; }
for k in 20 40 200; do run $k ""; run $k "--force-peer"; run $k "--force-peer --peer-blocking"; done
NFN_B200_PDL=0 run 20 "--force-peer"; NFN_B200_PDL=0 run 20 ""
