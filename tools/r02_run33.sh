#!/bin/bash
# round 2, last GPU seconds: the GPU arm's JSON line after the arm_config refactor (tiny rows, no extras)
cd "${GRAFT_REPO_ROOT:-.}"; mkdir -p gpurun_out; O=gpurun_out
timeout 25 python bench.py --rows 65536 --steps 3 --warmup 3 --no-cpu-baseline > $O/bench_cfgcheck.json 2> $O/bench_cfgcheck.err; echo "bench rc=$?"; cut -c1-900 $O/bench_cfgcheck.json; tail -n 5 $O/bench_cfgcheck.err | cut -c1-300
