#!/bin/bash
# round 2: fused Dense(P)+MDN kernel -- parity tests, timing vs the unfused composition, whole GPU suite
cd "${GRAFT_REPO_ROOT:-.}"; mkdir -p gpurun_out; O=gpurun_out
timeout 900 python -m pytest tests/test_dense_gpu.py -m gpu -x -q -k "mdn" > $O/pytest_dense_mdn.log 2>&1; echo "pytest mdn rc=$?"; tail -n 15 $O/pytest_dense_mdn.log
timeout 300 python tools/dense_mdn_time.py > $O/dense_mdn_time.txt 2>$O/dense_mdn_time.err; echo "time rc=$?"; cat $O/dense_mdn_time.txt; tail -n 5 $O/dense_mdn_time.err
timeout 1500 python -m pytest tests -m gpu -x -q > $O/pytest_gpu5.log 2>&1; echo "pytest all rc=$?"; tail -n 6 $O/pytest_gpu5.log
