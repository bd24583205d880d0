#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_layers_gpu.py tests/test_model_gpu.py tests/test_mbr_gpu.py -m gpu -x -q --timeout 300 2>&1 | tail -8 > gpurun_out/pytest_c.log
B=gpurun_out/bench_ab3.jsonl; : > $B
for cfg in "" "PK_GATE_BWD_TWO_PASS=1"; do
  echo "### ${cfg:-default}" >> $B
  env $cfg timeout 600 python bench.py --steps 8 --warmup 3 --no-cpu-baseline 2>/dev/null | tail -n 1 >> $B
done
timeout 300 python scripts/profile_step.py > gpurun_out/step_kernel_table.txt 2>&1
tail -n 4 gpurun_out/pytest_c.log; cut -c1-330 $B; grep -n "joint_gate\|total kernel\|lstm_seq" gpurun_out/step_kernel_table.txt | cut -c1-120
