#!/bin/bash
# what the driver runs at round end, on one GPU: GPU tests, smoke, default bench, reference arm; then the round-2
# ncu launch list of the bench command (per-launch durations, serialised: shares only)
cd "${GRAFT_REPO_ROOT:-.}"; mkdir -p gpurun_out; O=gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q > $O/pytest_final.log 2>&1; echo "pytest rc=$?"; tail -4 $O/pytest_final.log | cut -c1-200
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > $O/smoke_final.log 2>&1; echo "smoke rc=$?"; tail -2 $O/smoke_final.log
timeout 600 python bench.py --steps 20 --warmup 5 > $O/bench_final.json 2> $O/bench_final.err; echo "bench rc=$?"; cut -c1-200 $O/bench_final.json
timeout 600 python bench.py --impl reference --steps 20 --warmup 5 > $O/bench_ref_final.json 2>/dev/null; echo "ref rc=$?"; cut -c1-200 $O/bench_ref_final.json
timeout 400 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $O/r02_launches_cfg2.csv python bench.py --steps 5 --warmup 3 --no-cpu-baseline --no-other-configs > $O/ncu_launches.log 2>&1; echo "ncu launches rc=$?"
