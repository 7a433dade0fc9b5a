#!/usr/bin/env bash
# Build the experiment variants of libvae21.so that profiles/README.md refers to, into tools/ab/ (git-ignored; the directory
# still travels to the GPU box with gpurun).  Same flags as __graft_entry__.build() plus one compile-time switch each.
#   tools/build_variants.sh [timing] [mid] [xpf] [epi_single] ...     (default: timing mid xpf)
set -euo pipefail
cd "$(dirname "$0")/.."
mkdir -p tools/ab
NVCC=${NVCC:-$(command -v nvcc || echo /usr/local/cuda/bin/nvcc)}
FLAGS=(-gencode arch=compute_100a,code=sm_100a -lineinfo -O3 -std=c++17 -Xcompiler -fPIC -shared)
declare -A DEF=(
  [timing]="-DVAE21_TC_TIMING=1"          # per-role / per-record wait counters (tools/tc_timing.py)
  [mid]="-DVAE21_F32P_MID=1"              # FP32 kernel: hand-off bookkeeping in the middle of the stage body (rejected, +1.8 %)
  [xpf]="-DVAE21_F32P_XPF=1"              # FP32 kernel: software pipeline across the hand-off (rejected, +6.7 %)
  [epi_single]="-DVAE21_TC_EPI_SINGLE=1"  # TC kernel: epilogue without the two-register-set software pipeline
  [fp32_barrier]="-DVAE21_FP32_PIPE_DEFAULT=0"  # round-1 block-barrier FP32 kernel as the default
)
for v in "${@:-timing mid xpf}"; do
  for name in $v; do
    [[ -n "${DEF[$name]:-}" ]] || { echo "unknown variant $name (known: ${!DEF[*]})" >&2; exit 2; }
    echo "building tools/ab/libvae21_${name}.so (${DEF[$name]})"
    "$NVCC" "${FLAGS[@]}" ${DEF[$name]} -o "tools/ab/libvae21_${name}.so" 21cmvae_b200/csrc/vae21_api.cu
  done
done
