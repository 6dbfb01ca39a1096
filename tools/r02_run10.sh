#!/bin/bash
# round 2, tcgen05 fused Dense+chain kernel: pipe 2 (warp-specialised) vs pipe 1, correctness + timing
cd "${GRAFT_REPO_ROOT:-.}"; mkdir -p gpurun_out; O=gpurun_out
timeout 300 python tools/dense_time.py --steps 50 > $O/dense_time_pipe2.txt 2>$O/dense_time_pipe2.err; echo "pipe2 rc=$?"; cat $O/dense_time_pipe2.txt; tail -n 5 $O/dense_time_pipe2.err
NFN_B200_LIB=$PWD/normalizingflownetwork_b200/libnfn_b200_pipe1.so timeout 300 python tools/dense_time.py --steps 50 > $O/dense_time_pipe1.txt 2>$O/dense_time_pipe1.err; echo "pipe1 rc=$?"; cat $O/dense_time_pipe1.txt; tail -n 5 $O/dense_time_pipe1.err
timeout 900 python -m pytest tests/test_dense_gpu.py -m gpu -x -q > $O/pytest_dense.log 2>&1; echo "pytest dense rc=$?"; tail -n 8 $O/pytest_dense.log
