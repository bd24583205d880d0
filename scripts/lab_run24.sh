#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
O=gpurun_out/decode_gemm_ab.txt; : > $O
for cfg in "" "PK_GEMM_2SM=0" "PK_GEMM_2SM_MIN_TILES=100"; do
  env $cfg timeout 600 python bench.py --workload decode --steps 3 --warmup 3 --no-cpu-baseline 2>/dev/null | tail -n 1 | cut -c80-230 | sed "s/^/decode ${cfg:-default} /" >> $O
done
for cfg in "" "PK_GEMM_2SM_MIN_TILES=100"; do
  env $cfg timeout 600 python bench.py --steps 8 --warmup 3 --no-cpu-baseline 2>/dev/null | tail -n 1 | cut -c60-200 | sed "s/^/train ${cfg:-default} /" >> $O
done
cat $O
