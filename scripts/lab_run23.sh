#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
rm -f gpurun_out/parity_measured.jsonl
timeout 1500 python -m pytest tests -m gpu -q --timeout 300 2>&1 | tail -30 > gpurun_out/pytest_gpu.log
timeout 120 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" > gpurun_out/smoke.log 2>&1
timeout 300 python scripts/profile_step.py > gpurun_out/step_kernel_table.txt 2>&1
timeout 900 python bench.py --steps 10 --warmup 3 > gpurun_out/bench_train.json 2> gpurun_out/bench_train.err
timeout 600 python bench.py --workload decode --steps 3 --warmup 3 > gpurun_out/bench_decode.json 2> gpurun_out/bench_decode.err
timeout 600 python bench.py --workload mbr --steps 5 --warmup 3 > gpurun_out/bench_mbr.json 2> gpurun_out/bench_mbr.err
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 1500 --csv --log-file gpurun_out/r02_bench_launches.csv python bench.py --steps 2 --warmup 1 --no-cpu-baseline > gpurun_out/bench_under_ncu.log 2>&1
tail -n 4 gpurun_out/pytest_gpu.log; tail -n 1 gpurun_out/smoke.log; head -n 3 gpurun_out/step_kernel_table.txt | tail -n 1
python - <<'PY'
import json
for f in ["bench_train","bench_decode","bench_mbr"]:
    try:
        d=json.loads([l for l in open("gpurun_out/%s.json"%f) if l.startswith("{")][-1])
        print(f, {k:d.get(k) for k in ("value","ms_per_step","e2e","clocks")})
        r=d.get("roofline") or {}
        print("   roofline", {k:r.get(k) for k in ("achieved","peak","frac","launch_ms","alone")}); print("   cpu", (d.get("cpu_baseline") or {}).get("value"), "parity", d.get("parity_full_shape"))
    except Exception as e: print(f, "ERR", e)
PY
