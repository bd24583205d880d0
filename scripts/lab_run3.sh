#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
rm -f gpurun_out/parity_measured.jsonl
timeout 240 python -m pytest tests/test_layers_gpu.py -q -k "attention" 2>&1 | tail -60 > gpurun_out/pytest_attn.log
ATT=1; grep -q "failed\|Timeout\|error" gpurun_out/pytest_attn.log && ATT=0
[ -s gpurun_out/pytest_attn.log ] || ATT=0
echo "ATTN_TC usable: $ATT" >> gpurun_out/pytest_attn.log
PK_ATTN_TC=$ATT timeout 900 python -m pytest tests/test_model_gpu.py tests/test_rnnt_gpu.py tests/test_gemm_gpu.py tests/test_optim_gpu.py tests/test_decode_gpu.py tests/test_mbr_gpu.py -q --timeout 300 2>&1 | tail -80 > gpurun_out/pytest_gpu.log
B=gpurun_out/bench_ab.jsonl; : > $B
ab() { echo "### $*" >> $B; env "$@" timeout 400 python bench.py --steps 8 --warmup 3 --no-cpu-baseline 2>&1 | tail -2 >> $B; }
ab PK_ATTN_TC=$ATT
ab PK_ATTN_TC=$ATT PK_GEMM_SPLIT_MODE=1
ab PK_ATTN_TC=0
PK_ATTN_TC=$ATT PK_GEMM_SPLIT_MODE=1 timeout 300 python scripts/profile_step.py > gpurun_out/step_kernel_table.txt 2>&1
PK_ATTN_TC=$ATT timeout 300 python scripts/attn_bench.py > gpurun_out/attn_bench.txt 2>&1
tail -3 gpurun_out/pytest_attn.log; tail -5 gpurun_out/pytest_gpu.log; cat $B | cut -c1-330
