#!/bin/bash
# Round-2 GPU call 1: guard + tests + default bench + A/B of the two chain-kernel generations + ncu of a 2^18-row launch
cd "${GRAFT_REPO_ROOT:-.}"; mkdir -p gpurun_out; O=gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > $O/smi.txt 2>&1
timeout 300 python tools/tma_smoke.py > $O/tma_smoke.log 2>&1; SM=$?; echo "tma_smoke rc=$SM"; tail -3 $O/tma_smoke.log
if [ $SM -ne 0 ]; then SEL='-k not(tma)'; else SEL=''; fi
timeout 1500 python -m pytest tests -m gpu -x -q $SEL > $O/pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -5 $O/pytest_gpu.log
timeout 600 python bench.py --steps 20 --warmup 5 > $O/bench_default.json 2> $O/bench_default.err; echo "bench rc=$?"; cut -c1-600 $O/bench_default.json
ab() { # tag env cfg extra
  NFN_B200_CHAIN_IO=$2 timeout 300 python bench.py --config $3 --steps 100 --warmup 10 --no-cpu-baseline --no-other-configs $4 2>>$O/ab.err | python -c "
import json,sys
l=sys.stdin.readline()
try:
    j=json.loads(l); r=j['roofline']; print('$1'.ljust(16), '$2'.ljust(8), '$3', '$4'.ljust(12), 'ms/step %.4f frac %.3f value %.3e check %s'%(j['ms_per_step'], r['frac'], j['value'], (j.get('exchange_check') or {}).get('ok')))
except Exception as e: print('$1 $2 $3 $4 FAILED', e, l[:200])" | tee -a $O/ab.txt
}
for cfg in cfg2 cfg4; do
  ab base cpasync $cfg ""; ab base cpasync $cfg "--no-colsum"
  if [ $SM -eq 0 ]; then ab base tma $cfg ""; ab base tma $cfg "--no-colsum"; fi
done
ab base cpasync cfg3 ""; [ $SM -eq 0 ] && ab base tma cfg3 ""
ab base cpasync cfg2 "--fwd-only"; [ $SM -eq 0 ] && ab base tma cfg2 "--fwd-only"
if [ $SM -eq 0 ]; then
  for v in nb4 nb2w16 w12; do
    if [ -f normalizingflownetwork_b200/libnfn_b200_$v.so ]; then
      export NFN_B200_LIB=$PWD/normalizingflownetwork_b200/libnfn_b200_$v.so
      for cfg in cfg2 cfg4; do ab $v tma $cfg ""; done
      ab $v tma cfg3 ""
      unset NFN_B200_LIB
    fi
  done
fi
# row-count sweep (fixed vs marginal cost)
for rows in 262144 4194304; do
  ab base cpasync cfg2 "--rows $rows"; [ $SM -eq 0 ] && ab base tma cfg2 "--rows $rows"
  ab base cpasync cfg4 "--rows $rows"; [ $SM -eq 0 ] && ab base tma cfg4 "--rows $rows"
done
# ncu: one full capture of a 2^18-row launch per generation (plain run first, same command line)
for io in cpasync tma; do
  [ $io = tma ] && [ $SM -ne 0 ] && continue
  CMD="python bench.py --config cfg2 --rows 262144 --steps 5 --warmup 3 --no-cpu-baseline --no-other-configs"
  NFN_B200_CHAIN_IO=$io $CMD > $O/ncu_plain_$io.log 2>&1 && \
  NFN_B200_CHAIN_IO=$io timeout 600 ncu --set full --clock-control none --import-source on -k regex:chain_kernel -s 4 -c 2 -f -o $O/r02_cfg2_2e18_$io $CMD > $O/ncu_$io.log 2>&1
  echo "ncu $io rc=$?"
done
ls -la $O | tail -30
