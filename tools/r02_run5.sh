#!/bin/bash
# Round-2 GPU call 5 (2 GPUs): two-rank tests (peer exchange, data-parallel fit) + bench at N = 2
cd "${GRAFT_REPO_ROOT:-.}"; mkdir -p gpurun_out; O=gpurun_out
nvidia-smi -L > $O/smi_n2.txt
timeout 900 python -m pytest tests/test_peer_gpu.py tests/test_dp_fit_gpu.py -m gpu -x -q > $O/pytest_n2.log 2>&1; echo "pytest rc=$?"; tail -6 $O/pytest_n2.log
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511"
timeout 600 $TR bench.py --gpus 2 --steps 20 --warmup 5 > $O/bench_n2.json 2> $O/bench_n2.err; echo "bench n2 rc=$?"; cut -c1-250 $O/bench_n2.json
timeout 600 $TR bench.py --gpus 2 --steps 20 --warmup 5 --peer-blocking --no-other-configs > $O/bench_n2_blocking.json 2> $O/bench_n2_blocking.err; echo "rc=$?"; cut -c1-250 $O/bench_n2_blocking.json
timeout 600 $TR bench.py --gpus 2 --steps 20 --warmup 5 --exchange nccl --no-other-configs > $O/bench_n2_nccl.json 2> $O/bench_n2_nccl.err; echo "rc=$?"; cut -c1-250 $O/bench_n2_nccl.json
timeout 300 python bench.py --gpus 1 --steps 20 --warmup 5 --no-other-configs --no-cpu-baseline > $O/bench_n1_samebox.json 2>/dev/null; cut -c1-250 $O/bench_n1_samebox.json
