#!/bin/bash
# round 2, tcgen05 fused Dense+chain kernel with the h tiles staged by warps 5..7: correctness + timing + ncu
cd "${GRAFT_REPO_ROOT:-.}"; mkdir -p gpurun_out; O=gpurun_out
timeout 300 python tools/dense_time.py --steps 50 > $O/dense_time_stagers.txt 2>$O/dense_time_stagers.err; echo "stagers rc=$?"; cat $O/dense_time_stagers.txt; tail -n 5 $O/dense_time_stagers.err
timeout 900 python -m pytest tests/test_dense_gpu.py tests/test_estimators_gpu.py -m gpu -x -q > $O/pytest_dense.log 2>&1; echo "pytest dense rc=$?"; tail -n 4 $O/pytest_dense.log
CMD="python tools/dense_time.py --steps 3 --chains cfg2 --no-check"
timeout 600 ncu --set full --clock-control none --import-source on -k regex:dense_tc5_kernel -s 4 -c 1 -f -o $O/r02_tc5_stagers $CMD > $O/ncu_tc5_stagers.log 2>&1; echo "ncu rc=$?"
