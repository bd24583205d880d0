#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_decode_gpu.py tests/test_mbr_gpu.py tests/test_xf_prednet_gpu.py tests/test_trainer_cli_gpu.py -m gpu -x -q --timeout 300 2>&1 | tail -4 > gpurun_out/pytest_g.log
timeout 600 python scripts/profile_decode.py > gpurun_out/decode_kernel_table.txt 2>&1
for cfg in ""; do
  env $cfg timeout 600 python bench.py --workload decode --steps 3 --warmup 3 --no-cpu-baseline 2>/dev/null | tail -n 1 | cut -c1-220 | sed "s/^/${cfg:-default} /" >> gpurun_out/bench_decode_ab.txt
done
timeout 600 python bench.py --workload mbr --steps 5 --warmup 3 2>/dev/null | tail -n 1 | cut -c1-260 >> gpurun_out/bench_decode_ab.txt
tail -n 3 gpurun_out/pytest_g.log; sed -n 3,14p gpurun_out/decode_kernel_table.txt | cut -c1-150; cat gpurun_out/bench_decode_ab.txt
