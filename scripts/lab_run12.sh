#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 1200 python -m pytest tests/test_layers_gpu.py tests/test_model_gpu.py tests/test_decode_gpu.py tests/test_mbr_gpu.py tests/test_xf_prednet_gpu.py -m gpu -x -q --timeout 300 2>&1 | tail -30 > gpurun_out/pytest_b.log
B=gpurun_out/bench_ab2.jsonl; : > $B
for cfg in "" "PK_GATE_BWD_TWO_PASS=1"; do
  echo "### ${cfg:-default}" >> $B
  env $cfg timeout 600 python bench.py --steps 8 --warmup 3 --no-cpu-baseline 2>/dev/null | tail -n 1 >> $B
done
timeout 300 python scripts/profile_step.py > gpurun_out/step_kernel_table.txt 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:gemm_tcgen05 -s 1 -c 1 -o gpurun_out/r02_fc2_fwd_lse python scripts/gemm_lab.py lse > gpurun_out/ncu_lse.log 2>&1
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 1500 --csv --log-file gpurun_out/r02_bench_launches.csv python bench.py --steps 2 --warmup 1 --no-cpu-baseline > gpurun_out/bench_under_ncu.log 2>&1
tail -n 6 gpurun_out/pytest_b.log; cut -c1-330 $B; grep -n "joint_gate\|total kernel" gpurun_out/step_kernel_table.txt | cut -c1-120; tail -n 3 gpurun_out/ncu_lse.log; wc -l gpurun_out/r02_bench_launches.csv
