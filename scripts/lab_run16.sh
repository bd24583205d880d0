#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_layers_gpu.py tests/test_model_gpu.py tests/test_mbr_gpu.py tests/test_xf_prednet_gpu.py -m gpu -x -q --timeout 300 2>&1 | tail -5 > gpurun_out/pytest_d.log
for cfg in "" "PK_GATE_FWD_FRAME_MAJOR=1" "PK_GATE_BWD_OCC=3"; do
  echo "### ${cfg:-default}" > gpurun_out/table_$(echo ${cfg:-default} | tr '=' '_').txt
  env $cfg timeout 300 python scripts/profile_step.py 2>&1 | grep -E "total kernel|joint_gate|lstm_seq" | cut -c1-110 >> gpurun_out/table_$(echo ${cfg:-default} | tr '=' '_').txt
done
B=gpurun_out/bench_ab5.jsonl; : > $B
for cfg in "" "PK_GATE_FWD_FRAME_MAJOR=1 PK_GATE_BWD_TWO_PASS=1 PK_LSTM_CLUSTER=1 PK_LSTM_BARRIER=0"; do
  echo "### ${cfg:-default}" >> $B
  env $cfg timeout 600 python bench.py --steps 8 --warmup 3 --no-cpu-baseline 2>/dev/null | tail -n 1 >> $B
done
cat gpurun_out/pytest_d.log | tail -n 3; cat gpurun_out/table_*.txt; cut -c1-330 $B
