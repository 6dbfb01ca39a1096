#!/usr/bin/env bash
# Builds the C++ A/B harnesses into build/ (they travel to the GPU box with the snapshot):
#   build/dense_probe   fused Dense(P)+chain kernels through the C ABI: mma.sync vs tcgen05, no Python
#   build/umma_probe    tcgen05 descriptor semantics (K-/MN-major, no swizzle) and issue-rate microbenchmark
set -euo pipefail
cd "$(dirname "$0")/.."
NVCC=${NVCC:-/usr/local/cuda/bin/nvcc}
mkdir -p build
python -m normalizingflownetwork_b200.build
$NVCC -O2 -std=c++17 -o build/dense_probe tools/dense_probe.cu -Lnormalizingflownetwork_b200 -lnfn_b200 \
      -Xlinker -rpath -Xlinker '$ORIGIN/../normalizingflownetwork_b200' -Wno-deprecated-gpu-targets
$NVCC -O2 -std=c++17 -gencode arch=compute_100a,code=sm_100a -o build/umma_probe tools/umma_probe.cu
echo "built build/dense_probe build/umma_probe"
