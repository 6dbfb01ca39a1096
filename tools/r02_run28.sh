#!/bin/bash
# round 2: KMN bandwidth gradient as a block reduce-scatter -- mixture tests + timing
cd "${GRAFT_REPO_ROOT:-.}"; mkdir -p gpurun_out; O=gpurun_out
timeout 600 python -m pytest tests -m gpu -q -k "kmn or mixture or KMN or reference_code_run or estimators_match" > $O/pytest_kmn2.log 2>&1; echo "pytest kmn rc=$?"; tail -n 4 $O/pytest_kmn2.log | cut -c1-200
timeout 200 python tools/dense_kmn_time.py 2>&1 | grep -v Warn > $O/dense_kmn_time2.txt; cat $O/dense_kmn_time2.txt
