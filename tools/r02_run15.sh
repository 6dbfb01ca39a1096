#!/bin/bash
# round 2: folded-draw kernels (Bayesian step) -- tests, step profile, bench line
cd "${GRAFT_REPO_ROOT:-.}"; mkdir -p gpurun_out; O=gpurun_out
timeout 900 python -m pytest tests/test_draws_gpu.py -m gpu -x -q > $O/pytest_draws.log 2>&1; echo "pytest draws rc=$?"; tail -n 25 $O/pytest_draws.log
timeout 300 python tools/bayes_step_profile.py 2>&1 | grep -v Warn > $O/bayes_step_profile.txt; head -n 24 $O/bayes_step_profile.txt
timeout 1500 python -m pytest tests -m gpu -x -q > $O/pytest_gpu7.log 2>&1; echo "pytest all rc=$?"; tail -n 6 $O/pytest_gpu7.log
