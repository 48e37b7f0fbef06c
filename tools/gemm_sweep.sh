#!/bin/bash
# sweep the small-M GEMM knobs on the decode shapes (each config in its own process: the knobs are read once)
cd "$(dirname "$0")/.."
for cfg in "" "FIB200_GEMM_CLUSTER=1" "FIB200_GEMM_CLUSTER=4" "FIB200_GEMM_BM=128" "FIB200_GEMM_BM=128 FIB200_GEMM_CLUSTER=4" \
           "FIB200_GEMM_SWAP=1" "FIB200_GEMM_CLUSTER=2 FIB200_GEMM_BN=128" "FIB200_GEMM_CLUSTER=4 FIB200_GEMM_BN=128" \
           "FIB200_GEMM_CLUSTER=4 FIB200_GEMM_BN=256" "FIB200_GEMM_CLUSTER=2 FIB200_GEMM_BN=64" "FIB200_GEMM_CLUSTER=1 FIB200_GEMM_BN=32"; do
  echo "=== cfg: $cfg"
  env $cfg NOCUBLAS=1 SHAPES=qkv,o,down,gate_up timeout 120 python tools/gemm_decode_shapes.py 2>&1 | grep -v RESULT | sed 's/, .cublas_us.: 0.0//'
done
