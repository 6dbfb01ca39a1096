#!/bin/bash
# round 2: tensor-core hidden-layer backward, Bayesian MDN on the folded-draw kernels -- whole GPU suite (no -x), timings
cd "${GRAFT_REPO_ROOT:-.}"; mkdir -p gpurun_out; O=gpurun_out
timeout 1500 python -m pytest tests -m gpu -q > $O/pytest_gpu9.log 2>&1; echo "pytest all rc=$?"; tail -n 12 $O/pytest_gpu9.log | cut -c1-200
timeout 120 python tools/mlp_probe.py > $O/mlp_probe_mma.txt 2>&1; cat $O/mlp_probe_mma.txt | tail -n 3
NFN_B200_MLP_MMA=0 timeout 120 python tools/mlp_probe.py > $O/mlp_probe_scalar.txt 2>&1; cat $O/mlp_probe_scalar.txt | tail -n 3
timeout 200 python tools/estimator_breakdown.py > $O/estimator_breakdown2.log 2>&1; tail -n 12 $O/estimator_breakdown2.log | cut -c1-160
