#!/bin/bash
# round 2: per-draw weight-gradient pass retuned; ncu of the folded-draw Dense+chain kernel and the fused Dense+MDN kernel
cd "${GRAFT_REPO_ROOT:-.}"; mkdir -p gpurun_out; O=gpurun_out
timeout 600 python -m pytest tests/test_draws_gpu.py -m gpu -x -q > $O/pytest_draws2.log 2>&1; echo "pytest draws rc=$?"; tail -n 3 $O/pytest_draws2.log | cut -c1-200
timeout 300 python tools/bayes_step_profile.py 2>&1 | grep -v Warn > $O/bayes_step_profile4.txt; head -n 9 $O/bayes_step_profile4.txt | cut -c1-150
timeout 400 ncu --set full --clock-control none --import-source on -k regex:dense_chain_kernel -s 3 -c 1 -f -o $O/r02_dense_chain_draws python tools/bayes_step_profile.py > $O/ncu_draws.log 2>&1; echo "ncu draws rc=$?"
timeout 400 ncu --set full --clock-control none --import-source on -k regex:dense_mdn_kernel -s 4 -c 1 -f -o $O/r02_dense_mdn python tools/dense_mdn_time.py --steps 2 > $O/ncu_mdn.log 2>&1; echo "ncu mdn rc=$?"
