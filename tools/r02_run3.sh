#!/bin/bash
# Round-2 GPU call 3: full GPU tests (f3 fusion, new defaults) + fine geometry sweep + ncu of the new default
cd "${GRAFT_REPO_ROOT:-.}"; mkdir -p gpurun_out; O=gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q > $O/pytest_gpu3.log 2>&1; echo "pytest rc=$?"; tail -5 $O/pytest_gpu3.log
run() { # label cfg extra  (env passed by caller)
  timeout 200 python bench.py --config $2 --steps 100 --warmup 10 --no-cpu-baseline --no-other-configs --no-e2e $3 2>>$O/sweep3.err | python -c "
import json,sys
l=sys.stdin.readline()
try:
    j=json.loads(l); r=j['roofline']; print('$1'.ljust(26), '$2', '$3'.ljust(12), 'ms/step %.4f frac %.3f check %s'%(j['ms_per_step'], r['frac'], (j.get('exchange_check') or {}).get('ok')))
except Exception as e: print('$1 $2 $3 FAILED', e, l[:200])" | tee -a $O/sweep3.txt
}
for c in cfg2 cfg4 r10d1 r10d2; do run "aot-default" $c ""; NFN_B200_CHAIN_IO=cpasync run "aot-cpasync" $c ""; done
run "aot-default" cfg2 "--no-colsum"; run "aot-default" cfg3 ""; run "aot-default" cfg2 "--fwd-only"; run "aot-default" cfg4 "--fwd-only"
NFN_B200_CHAIN_IO=cpasync run "aot-cpasync" cfg4 "--fwd-only"
export NFN_B200_FORCE_JIT=1
t() { NFN_B200_TUNE_WNB=$1 NFN_B200_TUNE_WWARPS=$2 run "jit nb=$1 warps=$2" $3 "$4"; }
t 4 6 cfg2; t 4 7 cfg2; t 4 8 cfg2; t 4 9 cfg2; t 3 10 cfg2; t 3 11 cfg2; t 4 8 cfg2 --no-colsum
t 3 14 cfg4; t 3 16 cfg4; t 3 18 cfg4; t 4 14 cfg4
t 2 8 cfg4 --fwd-only; t 2 12 cfg4 --fwd-only; t 2 16 cfg4 --fwd-only; t 3 8 cfg4 --fwd-only
t 2 3 cfg3; t 2 4 cfg3; t 2 5 cfg3; t 3 3 cfg3
t 2 6 cfg2 --fwd-only; t 2 8 cfg2 --fwd-only; t 2 10 cfg2 --fwd-only; t 3 8 cfg2 --fwd-only; t 3 6 cfg2 --fwd-only
t 3 16 r10d1; t 4 12 r10d1; t 4 8 r10d1; t 3 12 r10d1
t 4 8 r10d2; t 3 12 r10d2; t 4 9 r10d2
unset NFN_B200_FORCE_JIT
timeout 600 python bench.py --steps 20 --warmup 5 > $O/bench_default3.json 2> $O/bench_default3.err; echo "bench rc=$?"; cut -c1-300 $O/bench_default3.json
CMD="python bench.py --config cfg2 --steps 5 --warmup 3 --no-cpu-baseline --no-other-configs --no-e2e"
$CMD > $O/ncu_plain3.log 2>&1 && timeout 600 ncu --set full --clock-control none --import-source on -k regex:chain_kernel -s 4 -c 2 -f -o $O/r02_cfg2_default $CMD > $O/ncu3.log 2>&1; echo "ncu rc=$?"
