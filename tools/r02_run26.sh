#!/bin/bash
# round 2: retuned folded first-layer kernels -- draws tests + profile
cd "${GRAFT_REPO_ROOT:-.}"; mkdir -p gpurun_out; O=gpurun_out
timeout 600 python -m pytest tests/test_draws_gpu.py -m gpu -q > $O/pytest_draws3.log 2>&1; echo "pytest draws rc=$?"; tail -n 3 $O/pytest_draws3.log | cut -c1-200
timeout 300 python tools/bayes_step_profile.py 2>&1 | grep -v Warn > $O/bayes_step_profile6.txt; head -n 8 $O/bayes_step_profile6.txt | cut -c1-150
