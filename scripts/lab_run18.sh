#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 600 python scripts/profile_decode.py > gpurun_out/decode_kernel_table.txt 2>&1
cat gpurun_out/decode_kernel_table.txt | cut -c1-160
