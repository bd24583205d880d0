#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
rm -f gpurun_out/parity_measured.jsonl
timeout 1500 python -m pytest tests -m gpu -q --timeout 300 2>&1 | tail -60 > gpurun_out/pytest_gpu.log
timeout 300 python scripts/profile_step.py > gpurun_out/step_kernel_table.txt 2>&1
timeout 900 python bench.py --steps 10 --warmup 3 > gpurun_out/bench_train.json 2> gpurun_out/bench_train.err
timeout 600 python bench.py --workload decode --steps 3 --warmup 3 > gpurun_out/bench_decode.json 2> gpurun_out/bench_decode.err
timeout 600 python bench.py --workload mbr --steps 5 --warmup 3 > gpurun_out/bench_mbr.json 2> gpurun_out/bench_mbr.err
timeout 600 python bench.py --batch 64 --steps 4 --warmup 3 --no-cpu-baseline > gpurun_out/bench_train_b64.json 2> gpurun_out/bench_train_b64.err
tail -4 gpurun_out/pytest_gpu.log; tail -32 gpurun_out/step_kernel_table.txt | cut -c1-150; cut -c1-400 gpurun_out/bench_train.json; cut -c1-300 gpurun_out/bench_decode.json gpurun_out/bench_mbr.json gpurun_out/bench_train_b64.json; tail -3 gpurun_out/*.err
