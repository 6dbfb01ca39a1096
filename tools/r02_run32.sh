#!/bin/bash
# round 2, last GPU seconds: config-2 chain density over a 2-D grid integrates to 1
cd "${GRAFT_REPO_ROOT:-.}"; mkdir -p gpurun_out; O=gpurun_out
timeout 40 python -m pytest tests/test_normalisation_gpu.py -q -k 2d > $O/pytest_norm2d.log 2>&1; echo "pytest norm2d rc=$?"; tail -n 30 $O/pytest_norm2d.log | cut -c1-600
