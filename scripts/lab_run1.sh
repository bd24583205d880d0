#!/bin/bash
# one gpurun call: GPU test suite, then the GEMM lab under several kernel-selection configurations
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
rm -f gpurun_out/parity_measured.jsonl
timeout 1200 python -m pytest tests -m gpu -q 2>&1 | tail -40 > gpurun_out/pytest_gpu.log
L=gpurun_out/gemm_lab.jsonl; : > $L
run() { echo "### $*" >> $L; env "$@" timeout 300 python scripts/gemm_lab.py $WHICH >> $L 2>&1; }
WHICH=all run PK_GEMM_2SM=0
WHICH=fc2 run PK_GEMM_2SM=0 PK_GEMM_L2_HINTS=0
WHICH=all run PK_GEMM_2SM=1 PK_GEMM_2SM_MIN_TILES=1
WHICH=fc2 run PK_GEMM_2SM=1 PK_GEMM_2SM_MIN_TILES=1 PK_GEMM_L2_HINTS=0
WHICH=wgrad run PK_GEMM_2SM=0 PK_GEMM_SPLIT_MODE=1
WHICH=wgrad run PK_GEMM_2SM=0 PK_GEMM_SPLIT_MODE=1 PK_GEMM_SPLIT_MAJOR=0
WHICH=wgrad run PK_GEMM_2SM=0 PK_GEMM_SPLIT_MAJOR=0
WHICH=wgrad run PK_GEMM_2SM=1 PK_GEMM_2SM_MIN_TILES=1 PK_GEMM_SPLIT_MODE=1
WHICH=wgrad run PK_GEMM_2SM=1 PK_GEMM_2SM_MIN_TILES=1 PK_GEMM_SPLIT_MODE=1 PK_GEMM_SPLIT_MAX=32
WHICH=enc run PK_GEMM_2SM=0 PK_GEMM_SPLIT_MODE=1
WHICH=enc run PK_GEMM_2SM=1 PK_GEMM_2SM_MIN_TILES=1 PK_GEMM_SPLIT_MODE=1
tail -5 gpurun_out/pytest_gpu.log
