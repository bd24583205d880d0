#!/bin/bash
# N GPUs of one box: NCCL BMUF trajectory test (2 ranks), then the torchrun bench lines at N = 2, 4, 8 (train) and N (mbr)
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
N=${1:-2}
timeout 400 python -m pytest tests/test_optim_gpu.py -m gpu -q --timeout 300 -k nccl 2>&1 | tail -5 > gpurun_out/pytest_nccl.log
for n in ${TRAIN_NS:-2 4 8}; do
  [ $n -le $N ] || continue
  timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port $((29600+n)) \
      bench.py --gpus $n --steps 10 --warmup 3 > gpurun_out/bench_train_N$n.json 2> gpurun_out/bench_train_N$n.err
done
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29700 \
    bench.py --gpus $N --workload mbr --steps 5 --warmup 3 > gpurun_out/bench_mbr_N$N.json 2> gpurun_out/bench_mbr_N$N.err
[ -n "$SKIP_DECODE" ] || timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29710 \
    bench.py --gpus $N --workload decode --steps 3 --warmup 3 > gpurun_out/bench_decode_N$N.json 2> gpurun_out/bench_decode_N$N.err
cat gpurun_out/pytest_nccl.log; for f in gpurun_out/bench_*_N*.json; do echo $f; cut -c1-330 $f; done; for f in gpurun_out/bench_*_N*.err; do tail -n 2 $f; done
