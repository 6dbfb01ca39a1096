#!/bin/bash
# round 2, last GPU seconds: canary (guard-band) test of the round-2 kernels' output writes
cd "${GRAFT_REPO_ROOT:-.}"; mkdir -p gpurun_out; O=gpurun_out
timeout 95 python -m pytest tests/test_guard_bands_gpu.py -x -q > $O/pytest_guard.log 2>&1; echo "pytest guard rc=$?"; tail -n 40 $O/pytest_guard.log | cut -c1-400
