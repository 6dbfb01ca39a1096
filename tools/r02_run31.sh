#!/bin/bash
# round 2, last GPU seconds: the CUDA heads' densities integrate to 1 (first-principles check, no oracle)
cd "${GRAFT_REPO_ROOT:-.}"; mkdir -p gpurun_out; O=gpurun_out
timeout 70 python -m pytest tests/test_normalisation_gpu.py -q > $O/pytest_norm.log 2>&1; echo "pytest norm rc=$?"; tail -n 40 $O/pytest_norm.log | cut -c1-600
