#!/bin/bash
# Round-2 GPU call 4: alignment probe, sigma = 1 tail, forward small-tile sweep, GPU tests
cd "${GRAFT_REPO_ROOT:-.}"; mkdir -p gpurun_out; O=gpurun_out
timeout 300 python tools/align_probe.py cfg2 > $O/align_cfg2.txt 2>&1; tail -20 $O/align_cfg2.txt
timeout 300 python tools/align_probe.py r10d2 > $O/align_r10d2.txt 2>&1; tail -8 $O/align_r10d2.txt
timeout 600 python tools/sigma1_tail.py > $O/sigma1_tail.md 2>&1; cat $O/sigma1_tail.md
run() { # label cfg extra  (env passed by caller)
  timeout 200 python bench.py --config $2 --steps 100 --warmup 10 --no-cpu-baseline --no-other-configs --no-e2e $3 2>>$O/sweep4.err | python -c "
import json,sys
l=sys.stdin.readline()
try:
    j=json.loads(l); r=j['roofline']; print('$1'.ljust(26), '$2', '$3'.ljust(12), 'ms/step %.4f frac %.3f check %s'%(j['ms_per_step'], r['frac'], (j.get('exchange_check') or {}).get('ok')))
except Exception as e: print('$1 $2 $3 FAILED', e, l[:200])" | tee -a $O/sweep4.txt
}
export NFN_B200_FORCE_JIT=1
t() { NFN_B200_TUNE_WNB=$1 NFN_B200_TUNE_WWARPS=$2 run "jit nb=$1 warps=$2" $3 "$4"; }
t 2 24 cfg4 --fwd-only; t 2 32 cfg4 --fwd-only; t 3 16 cfg4 --fwd-only; t 4 12 cfg4 --fwd-only; t 4 16 cfg4 --fwd-only; t 3 24 cfg4 --fwd-only; t 3 32 cfg4 --fwd-only
t 2 16 r10d1 --fwd-only; t 2 8 r10d1 --fwd-only; t 3 12 r10d1 --fwd-only; t 3 16 r10d1 --fwd-only
t 4 8 r10d2; t 4 8 r10d2; t 3 12 r10d2
unset NFN_B200_FORCE_JIT
run "aot-default" r10d2 ""; run "aot-default" r10d2 ""; run "aot-default" r10d1 "--fwd-only"
timeout 1500 python -m pytest tests -m gpu -x -q > $O/pytest_gpu4.log 2>&1; echo "pytest rc=$?"; tail -5 $O/pytest_gpu4.log
