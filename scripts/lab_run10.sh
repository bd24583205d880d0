#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_xf_prednet_gpu.py tests/test_layers_gpu.py tests/test_decode_gpu.py -m gpu -x -q --timeout 300 2>&1 | tail -40 > gpurun_out/pytest_xf.log
timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/bench_train_host.json 2> gpurun_out/bench_train_host.err
tail -n 25 gpurun_out/pytest_xf.log; python - <<'PY'
import json
d=json.loads([l for l in open("gpurun_out/bench_train_host.json") if l.startswith("{")][-1])
print({k:d.get(k) for k in ("value","ms_per_step","host_ms_per_step","clocks","gpu_launches")})
PY
