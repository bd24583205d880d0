#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
rm -f gpurun_out/parity_measured.jsonl
timeout 600 python -m pytest tests/test_gemm_gpu.py tests/test_layers_gpu.py tests/test_frontend_gpu.py -q --timeout 300 2>&1 | tail -30 > gpurun_out/pytest_attn.log
timeout 300 python scripts/attn_bench.py > gpurun_out/attn_bench.txt 2>&1
L=gpurun_out/gemm_lab3.jsonl; : > $L
echo "### default" >> $L; timeout 300 python scripts/gemm_lab.py fc2 >> $L 2>&1
B=gpurun_out/bench_ab.jsonl; : > $B
echo "### default" >> $B; timeout 400 python bench.py --steps 8 --warmup 3 --no-cpu-baseline 2>&1 | tail -2 >> $B
echo "### PK_ATTN_TC=0" >> $B; PK_ATTN_TC=0 timeout 400 python bench.py --steps 8 --warmup 3 --no-cpu-baseline 2>&1 | tail -2 >> $B
timeout 300 python scripts/profile_step.py > gpurun_out/step_kernel_table.txt 2>&1
tail -4 gpurun_out/pytest_attn.log; cat gpurun_out/attn_bench.txt; grep shape $L | cut -c1-160; cut -c1-300 $B
