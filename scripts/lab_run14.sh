#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
L=gpurun_out/lstm_bench.txt; : > $L
timeout 300 python -m pytest tests/test_layers_gpu.py -m gpu -x -q --timeout 120 -k "lstm" 2>&1 | tail -5 >> $L
for cl in 1 8 4; do for bar in 0 1 2; do
  PK_LSTM_CLUSTER=$cl PK_LSTM_BARRIER=$bar timeout 120 python scripts/lstm_bench.py 2>&1 | tail -n 1 | sed "s/^/cluster=$cl /" >> $L
done; done
B=gpurun_out/bench_ab4.jsonl; : > $B
for cfg in "" "PK_LSTM_CLUSTER=1 PK_LSTM_BARRIER=0"; do
  echo "### ${cfg:-default}" >> $B
  env $cfg timeout 600 python bench.py --steps 8 --warmup 3 --no-cpu-baseline 2>/dev/null | tail -n 1 >> $B
done
cat $L; cut -c1-330 $B
