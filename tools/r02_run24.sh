#!/bin/bash
# round 2: fused Dense(P)+KMN head -- dense / mixture tests, then the whole suite
cd "${GRAFT_REPO_ROOT:-.}"; mkdir -p gpurun_out; O=gpurun_out
timeout 900 python -m pytest tests/test_dense_gpu.py tests/test_parity_gpu.py -m gpu -q -k "kmn or mdn or mixture" > $O/pytest_kmn.log 2>&1; echo "pytest kmn rc=$?"; tail -n 15 $O/pytest_kmn.log | cut -c1-220
timeout 1500 python -m pytest tests -m gpu -q > $O/pytest_gpu11.log 2>&1; echo "pytest all rc=$?"; tail -n 8 $O/pytest_gpu11.log | cut -c1-200
