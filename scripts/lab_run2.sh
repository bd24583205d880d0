#!/bin/bash
# gpurun call 3: attention (tcgen05) tests first under a short timeout, full GPU suite, step-level A/B, kernel table, wgrad sweep, ncu of the joint forward
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
rm -f gpurun_out/parity_measured.jsonl
timeout 240 python -m pytest tests/test_layers_gpu.py -q -k "attention" 2>&1 | tail -40 > gpurun_out/pytest_attn.log
ATT=1; grep -q "failed\|Timeout\|error" gpurun_out/pytest_attn.log && ATT=0
[ -s gpurun_out/pytest_attn.log ] || ATT=0
echo "ATTN_TC usable: $ATT" >> gpurun_out/pytest_attn.log
PK_ATTN_TC=$ATT timeout 1200 python -m pytest tests -m gpu -q --timeout 300 2>&1 | tail -150 > gpurun_out/pytest_gpu.log
B=gpurun_out/bench_ab.jsonl; : > $B
ab() { echo "### $*" >> $B; env "$@" timeout 400 python bench.py --steps 8 --warmup 3 --no-cpu-baseline 2>&1 | tail -2 >> $B; }
ab PK_ATTN_TC=$ATT
ab PK_ATTN_TC=0 PK_GEMM_2SM=0
ab PK_ATTN_TC=0
[ "$ATT" = 1 ] && ab PK_ATTN_TC=1 PK_FOLD_MASKS=0
PK_ATTN_TC=$ATT timeout 300 python scripts/profile_step.py > gpurun_out/step_kernel_table.txt 2>&1
L=gpurun_out/gemm_lab2.jsonl; : > $L
run() { echo "### $*" >> $L; env "$@" timeout 300 python scripts/gemm_lab.py $WHICH >> $L 2>&1; }
WHICH=wgrad run PK_GEMM_2SM=1
WHICH=wgrad run PK_GEMM_2SM=1 PK_GEMM_SPLIT_MAJOR=0
WHICH=wgrad run PK_GEMM_2SM=1 PK_GEMM_SPLIT_MODE=1
WHICH=wgrad run PK_GEMM_2SM=1 PK_GEMM_SPLIT_MODE=1 PK_GEMM_SPLIT_MAX=32
WHICH=wgrad run PK_GEMM_2SM=0
WHICH=fc2 run PK_GEMM_2SM=1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:gemm_tcgen05 -c 2 -o gpurun_out/r02_fc2_fwd python scripts/gemm_lab.py fwd > gpurun_out/ncu_fwd.log 2>&1
tail -3 gpurun_out/pytest_attn.log; tail -5 gpurun_out/pytest_gpu.log; cat $B | cut -c1-400
