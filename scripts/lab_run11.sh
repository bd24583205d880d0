#!/bin/bash
# 8-GPU diagnosis: per-rank step times, BMUF sync time, per-GPU clocks / power during the run
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
nvidia-smi --query-gpu=index,clocks.sm,power.draw,temperature.gpu,clocks_event_reasons.sw_power_cap,clocks_event_reasons.sw_thermal_slowdown --format=csv,noheader,nounits -lms 250 > gpurun_out/smi_n8.csv 2>/dev/null &
SMI=$!
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29608 \
    bench.py --gpus 8 --steps 20 --warmup 6 --no-cpu-baseline > gpurun_out/bench_train_N8b.json 2> gpurun_out/bench_train_N8b.err
kill $SMI
python - <<'PY'
import json, collections
d=json.loads([l for l in open("gpurun_out/bench_train_N8b.json") if l.startswith("{")][-1])
print({k:d.get(k) for k in ("value","ms_per_step","host_ms_per_step","per_rank_ms_per_step","bmuf_sync_ms","clocks")}, d["e2e"])
rows=collections.defaultdict(list)
for l in open("gpurun_out/smi_n8.csv"):
    f=[x.strip() for x in l.split(",")]
    if len(f)>=4:
        try: rows[f[0]].append((float(f[1]),float(f[2]),float(f[3])))
        except ValueError: pass
for g,v in sorted(rows.items()):
    busy=[r for r in v if r[1]>400]
    if busy:
        sm=sorted(r[0] for r in busy); pw=sorted(r[1] for r in busy); tp=max(r[2] for r in busy)
        print("gpu",g,"busy samples",len(busy),"sm median",sm[len(sm)//2],"min",sm[0],"power median",pw[len(pw)//2],"max",pw[-1],"temp max",tp)
PY
