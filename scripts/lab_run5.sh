#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
rm -f gpurun_out/parity_measured.jsonl
timeout 240 python -m pytest tests/test_layers_gpu.py -q -k "attention" 2>&1 | tail -30 > gpurun_out/pytest_attn.log
ATT=1; grep -q "failed\|Timeout\|error" gpurun_out/pytest_attn.log && ATT=0
[ -s gpurun_out/pytest_attn.log ] || ATT=0
echo "ATTN_TC usable: $ATT" >> gpurun_out/pytest_attn.log
PK_ATTN_TC=$ATT timeout 1200 python -m pytest tests -m gpu -q --timeout 300 2>&1 | tail -80 > gpurun_out/pytest_gpu.log
PK_ATTN_TC=$ATT timeout 300 python scripts/attn_bench.py > gpurun_out/attn_bench.txt 2>&1
B=gpurun_out/bench_ab.jsonl; : > $B
ab() { echo "### $*" >> $B; env "$@" timeout 400 python bench.py --steps 8 --warmup 3 --no-cpu-baseline 2>&1 | tail -2 >> $B; }
ab PK_ATTN_TC=$ATT PK_GEMM_SPLIT_MODE=1
ab PK_ATTN_TC=$ATT
echo "### decode" >> $B; timeout 600 python bench.py --workload decode --steps 2 --warmup 3 --no-cpu-baseline 2>&1 | tail -3 >> $B
echo "### mbr" >> $B; timeout 600 python bench.py --workload mbr --steps 3 --warmup 3 2>&1 | tail -3 >> $B
PK_ATTN_TC=$ATT PK_GEMM_SPLIT_MODE=1 timeout 300 python scripts/profile_step.py > gpurun_out/step_kernel_table.txt 2>&1
tail -3 gpurun_out/pytest_attn.log; tail -6 gpurun_out/pytest_gpu.log; cat gpurun_out/attn_bench.txt; cat $B | cut -c1-330
