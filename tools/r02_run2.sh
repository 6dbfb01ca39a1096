#!/bin/bash
# Round-2 GPU call 2: full GPU tests + geometry sweep of the warp-tile kernels through the runtime specialiser
cd "${GRAFT_REPO_ROOT:-.}"; mkdir -p gpurun_out; O=gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q > $O/pytest_gpu2.log 2>&1; echo "pytest rc=$?"; tail -5 $O/pytest_gpu2.log
run() { # label cfg extra  (env passed by caller)
  timeout 200 python bench.py --config $2 --steps 100 --warmup 10 --no-cpu-baseline --no-other-configs --no-e2e $3 2>>$O/sweep.err | python -c "
import json,sys
l=sys.stdin.readline()
try:
    j=json.loads(l); r=j['roofline']; print('$1'.ljust(26), '$2', '$3'.ljust(12), 'ms/step %.4f frac %.3f check %s'%(j['ms_per_step'], r['frac'], (j.get('exchange_check') or {}).get('ok')))
except Exception as e: print('$1 $2 $3 FAILED', e, l[:200])" | tee -a $O/sweep.txt
}
export NFN_B200_CHAIN_IO=tma
run "aot-default" cfg2 ""; run "aot-default" cfg2 "--no-colsum"; run "aot-default" cfg4 ""; run "aot-default" cfg4 "--no-colsum"; run "aot-default" cfg3 ""; run "aot-default" cfg2 "--fwd-only"
export NFN_B200_FORCE_JIT=1
for nb in 2 3; do for w in 8 12 16; do
  NFN_B200_TUNE_WNB=$nb NFN_B200_TUNE_WWARPS=$w run "jit nb=$nb warps=$w" cfg2 ""
done; done
NFN_B200_TUNE_WNB=4 NFN_B200_TUNE_WWARPS=8 run "jit nb=4 warps=8" cfg2 ""
NFN_B200_TUNE_WNB=2 NFN_B200_TUNE_WWARPS=16 run "jit nb=2 warps=16" cfg2 "--no-colsum"
for nb in 2 3 4; do for w in 8 12 16 20 24; do
  NFN_B200_TUNE_WNB=$nb NFN_B200_TUNE_WWARPS=$w run "jit nb=$nb warps=$w" cfg4 ""
done; done
NFN_B200_TUNE_WNB=3 NFN_B200_TUNE_WWARPS=16 run "jit nb=3 warps=16" cfg4 "--no-colsum"
for nb in 2 3; do for w in 4 6; do
  NFN_B200_TUNE_WNB=$nb NFN_B200_TUNE_WWARPS=$w run "jit nb=$nb warps=$w" cfg3 ""
done; done
for w in 8 12 16 20; do NFN_B200_TUNE_WNB=2 NFN_B200_TUNE_WWARPS=$w run "jit nb=2 warps=$w" cfg2 "--fwd-only"; done
unset NFN_B200_FORCE_JIT NFN_B200_CHAIN_IO
timeout 600 python bench.py --steps 20 --warmup 5 > $O/bench_default2.json 2> $O/bench_default2.err; echo "bench rc=$?"; cut -c1-400 $O/bench_default2.json
