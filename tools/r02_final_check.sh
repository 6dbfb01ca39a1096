#!/bin/bash
# what the driver runs at round end, on one GPU: GPU tests, smoke, default bench, reference arm
cd "${GRAFT_REPO_ROOT:-.}"; mkdir -p gpurun_out; O=gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q > $O/pytest_final.log 2>&1; echo "pytest rc=$?"; tail -4 $O/pytest_final.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > $O/smoke_final.log 2>&1; echo "smoke rc=$?"; tail -2 $O/smoke_final.log
timeout 600 python bench.py --steps 20 --warmup 5 > $O/bench_final.json 2> $O/bench_final.err; echo "bench rc=$?"; cut -c1-200 $O/bench_final.json
timeout 600 python bench.py --impl reference --steps 20 --warmup 5 > $O/bench_ref_final.json 2>/dev/null; echo "ref rc=$?"; cut -c1-200 $O/bench_ref_final.json
