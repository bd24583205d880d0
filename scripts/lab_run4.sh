#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
B=8 timeout 600 ncu --set full --clock-control none --import-source on -k regex:attention_tc -s 12 -c 3 -o gpurun_out/r02_attn_tc python scripts/attn_bench.py > gpurun_out/ncu_attn.log 2>&1
tail -5 gpurun_out/ncu_attn.log
