#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
PK_DECODE_GRAPH=0 timeout 600 ncu --set full --clock-control none --import-source on -k regex:beam_advance -s 200 -c 1 -o gpurun_out/r02_beam_advance python scripts/profile_decode.py > gpurun_out/ncu_adv.log 2>&1
tail -n 3 gpurun_out/ncu_adv.log
