#!/bin/bash
# Build libmipb200 from the sources of a git revision into vvc-mip-gpu_b200/lib/<name>.so for A/B timing with
# tools/gpu_sweep.sh (MIPB200_LIB).  Usage: tools/build_variant.sh <git-rev> <name>
set -eu
rev=$1; name=$2
root=$(git rev-parse --show-toplevel)
tmp=$(mktemp -d)
git -C "$root" archive "$rev" vvc-mip-gpu_b200/csrc include | tar -x -C "$tmp"
mkdir -p "$root/vvc-mip-gpu_b200/lib"
( cd "$tmp/vvc-mip-gpu_b200/csrc" && /usr/local/cuda/bin/nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -lineinfo \
    -Xcompiler -fPIC,-fvisibility=hidden -ccbin g++ -shared -o "$root/vvc-mip-gpu_b200/lib/$name.so" mip_kernels.cu mip_engine.cu )
rm -rf "$tmp"
echo "built vvc-mip-gpu_b200/lib/$name.so from $rev"
