#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
L=gpurun_out/lstm_bench2.txt; : > $L
timeout 300 python -m pytest tests/test_layers_gpu.py -m gpu -x -q --timeout 120 -k "lstm" 2>&1 | tail -3 >> $L
for bar in 1 3; do
  PK_LSTM_CLUSTER=8 PK_LSTM_BARRIER=$bar timeout 120 python scripts/lstm_bench.py 2>&1 | tail -n 1 | sed "s/^/cluster=8 /" >> $L
done
B=64 PK_LSTM_CLUSTER=8 PK_LSTM_BARRIER=3 timeout 120 python scripts/lstm_bench.py 2>&1 | tail -n 1 | sed "s/^/B=64 cluster=8 /" >> $L
cat $L
