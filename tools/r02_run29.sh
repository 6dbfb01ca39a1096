#!/bin/bash
# round 2, last check of the final tree: whole GPU suite + smoke
cd "${GRAFT_REPO_ROOT:-.}"; mkdir -p gpurun_out; O=gpurun_out
timeout 600 python -m pytest tests -m gpu -x -q > $O/pytest_final2.log 2>&1; echo "pytest rc=$?"; tail -3 $O/pytest_final2.log | cut -c1-200
timeout 120 python -c "import __graft_entry__ as g; g.smoke()" > $O/smoke_final2.log 2>&1; echo "smoke rc=$?"; tail -1 $O/smoke_final2.log
