#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_decode_gpu.py tests/test_mbr_gpu.py tests/test_xf_prednet_gpu.py -m gpu -x -q --timeout 300 2>&1 | tail -4 > gpurun_out/pytest_f.log
timeout 600 python scripts/profile_decode.py > gpurun_out/decode_kernel_table.txt 2>&1
timeout 600 python bench.py --workload decode --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/bench_decode3.json 2> gpurun_out/bench_decode3.err
tail -n 3 gpurun_out/pytest_f.log; sed -n 3,14p gpurun_out/decode_kernel_table.txt | cut -c1-150
python - <<'PY'
import json
for f in ["bench_decode3"]:
    d=json.loads([l for l in open("gpurun_out/%s.json"%f) if l.startswith("{")][-1])
    print(f, {k:d.get(k) for k in ("value","ms_per_step","e2e","gpu_launches")})
PY
