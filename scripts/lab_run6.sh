#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_layers_gpu.py tests/test_frontend_gpu.py -q -k "attention or dither or lstm" 2>&1 | tail -30 > gpurun_out/pytest_attn.log
timeout 300 python scripts/attn_bench.py > gpurun_out/attn_bench.txt 2>&1
B=8 timeout 600 ncu --set full --clock-control none --import-source on -k regex:attention_tc -s 12 -c 3 -o gpurun_out/r02_attn_tc_v4 python scripts/attn_bench.py > gpurun_out/ncu_attn.log 2>&1
tail -4 gpurun_out/pytest_attn.log; cat gpurun_out/attn_bench.txt
