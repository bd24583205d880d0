#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
B=8 timeout 600 ncu --set full --clock-control none --import-source on -k regex:attention_tc_kernel -s 12 -c 3 -o gpurun_out/r02_attn_tc_v5 python scripts/attn_bench.py > gpurun_out/ncu_attn5.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:joint_gate -s 2 -c 2 -o gpurun_out/r02_joint_gate python bench.py --steps 1 --warmup 1 --no-cpu-baseline > gpurun_out/ncu_gate.log 2>&1
tail -n 2 gpurun_out/ncu_attn5.log gpurun_out/ncu_gate.log
