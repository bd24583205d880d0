#!/usr/bin/env python
"""Benchmark of the B200-native RNN-T train step (BASELINE.json metric:
"utterances/sec RNN-T train step (T=1000,U=150,V=6k)"), configs[1]: batch 32 per GPU, bf16.

    python bench.py --gpus N --steps K --warmup W            # ours (one process per GPU under torchrun)
    python bench.py --impl reference --gpus N --steps K ...  # the CPU arm (oracle port of the reference path)

A step = one pass of the hot path over one synthetic batch: H2D of raw 16 kHz PCM (e2e only) ->
on-GPU speed/gain augmentation + fbank + splice + CMN/CMVN + SpecAugment -> TDNN-Transformer encoder,
LSTM prediction net, fused joint + RNN-T loss, full backward -> inf-norm clip + Nesterov SGD ->
BMUF block sync every 5th step (NCCL all-reduce when N > 1).  Prints ONE JSON line on rank 0.
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time
import types

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

METRIC = "utterances/sec RNN-T train step (T=1000,U=150,V=6k)"


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference", "_cpu_worker"])
    ap.add_argument("--batch", type=int, default=32)
    ap.add_argument("--T", type=int, default=1000)
    ap.add_argument("--U", type=int, default=150)
    ap.add_argument("--V", type=int, default=6000)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--cpu-budget-s", type=float, default=150.0)
    return ap.parse_args()


def model_args(V):
    return types.SimpleNamespace(rnn_size=1024, local_rank=0, decoder_type="rnn", brnn=True, encoder_type="transformer",
                                 embd_dim=100, padding_idx=V, dropout=0.2, dec_layers=2, enc_layers=9)


def train_args():
    # egs/train_transducer_bmuf_otfaug.sh:157-197 (recipe values); SpecAugment on (north star)
    return types.SimpleNamespace(cmn=True, model_lctx=21, model_rctx=21, model_stride=4, sync_period=5, initial_lr=4e-4,
                                 final_lr=4e-5, momentum=0.9, grad_clip=3.0, num_epochs=15, num_batches_per_epoch=1000, epoch=0,
                                 block_momentum=0.9, block_lr=1.0, max_freq_span=15, max_time_span=35)


def synth_pcm(B, T, seed):
    import numpy as np
    n = 400 + (T - 1) * 160
    rng = np.random.default_rng(seed)
    return np.clip(np.round(rng.normal(0.0, 3000.0, (B, n))), -32768, 32767).astype(np.int16)


def peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        d = json.load(open(p))
        return d["hbm_gbs"], d["bf16_tflops"], d.get("bf16_tflops_sustained", d["bf16_tflops"]), "measured"
    return 6650.0, 1590.0, 1400.0, "fallback"


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled DURING the timed region (B200_PROFILING.md)."""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index):
        self.gpu, self.rows, self.proc = gpu_index, [], None

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "--query-gpu=" + self.Q, "--format=csv,noheader,nounits", "-lms", "200"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            threading.Thread(target=self._read, daemon=True).start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            f = [x.strip() for x in line.split(",")]
            if len(f) >= 8 and f[0] == str(self.gpu):
                self.rows.append(f)

    def stop(self):
        if self.proc is not None:
            self.proc.kill()
        sm = sorted(int(float(r[1])) for r in self.rows if r[1].replace(".", "").isdigit())
        mx = [int(float(r[2])) for r in self.rows if r[2].replace(".", "").isdigit()]
        reasons = set()
        for r in self.rows:
            for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), r[4:8]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": max(mx) if mx else None, "reasons": sorted(reasons),
                "samples": len(self.rows)}


# ------------------------------------------------------------------------------------------------ CPU arm
NCU_FC2_TRAFFIC_BYTES = 17.231e9       # dram__bytes_read.sum + dram__bytes_write.sum of one plain fc2 launch (profiles/r01_gemm_fc2.ncu.txt)
NCU_FC2_LSE_TRAFFIC_BYTES = 18.768e9   # same for the launch with the row-LSE epilogue (profiles/r01_gemm_fc2_lse.ncu.txt)
CPU_THREADS_CAP = 32      # torch-CPU fp32 layers stop scaling (and oversubscribe) beyond a few dozen threads on the 128-core hosts


def cpu_step_fn(a, seed=777):
    """Builds the oracle port of one training batch at B=1 (the bounded sample) and returns (fn, cores)."""
    import numpy as np
    import torch
    from oracle import train_step as ots
    from pika_b200.model.transducer import Net
    cores = min(os.cpu_count() or 1, CPU_THREADS_CAP)
    torch.set_num_threads(cores)
    torch.manual_seed(seed)
    net = Net(model_args(a.V), 240, a.V)                     # parameter container only (CPU); weights = reference init
    sd = {k: v.detach().clone() for k, v in net.state_dict().items()}
    for k, _ in net.named_parameters():
        sd[k].requires_grad_(True)
    pcm = synth_pcm(1, a.T, seed)
    rng = np.random.default_rng(seed)
    labels = rng.integers(1, a.V, (1, a.U)).astype(np.int32)
    ta = train_args()
    state = {"bufs": None}

    def fn():
        data, lens = ots.features([pcm[0]], [1.0], [-25.0], cmn=True, specaug=(100, 7, 300, 20))
        tl = (lens - ta.model_lctx - ta.model_rctx)
        tl = tl // ta.model_stride + (tl % ta.model_stride != 0)
        costs, state["bufs"] = ots.train_step(sd, data, labels, tl.astype(np.int32), np.array([a.U], np.int32), ta.initial_lr,
                                              ta.momentum, ta.grad_clip, state["bufs"])
        return float(costs.sum())
    return fn, cores


def _cpu_worker(argv):
    """subprocess entry: times `steps` CPU batches after `warm` warm-ups and prints one JSON line"""
    a = parse_from(argv)
    fn, cores = cpu_step_fn(a)
    t0 = time.time()
    first = None
    done = 0
    for i in range(a.warmup + a.steps):
        t1 = time.time()
        fn()
        dt1 = time.time() - t1
        if first is None:
            first = dt1
        if i >= a.warmup:
            done += 1
        print(json.dumps({"progress": i + 1, "step_s": dt1}), flush=True)
    print(json.dumps({"done": done, "first_s": first, "total_s": time.time() - t0, "cores": cores}), flush=True)


def parse_from(argv):
    old = sys.argv
    sys.argv = [old[0]] + list(argv)
    try:
        return parse()
    finally:
        sys.argv = old


def run_cpu_bounded(a, warmup, steps, budget_s):
    """Runs the CPU arm in a subprocess with a hard wall-clock budget; returns (mean step seconds over the timed
    steps that finished, number of timed steps, number of warm-ups, cores)."""
    cmd = [sys.executable, os.path.abspath(__file__), "--impl", "_cpu_worker", "--T", str(a.T), "--U", str(a.U), "--V", str(a.V),
           "--steps", str(steps), "--warmup", str(warmup)]
    p = subprocess.Popen(cmd, stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
    t0 = time.time()
    times, cores = [], min(os.cpu_count() or 1, CPU_THREADS_CAP)
    import selectors
    sel = selectors.DefaultSelector()
    sel.register(p.stdout, selectors.EVENT_READ)
    while True:
        left = budget_s - (time.time() - t0)
        if left <= 0:
            p.kill()
            break
        if not sel.select(timeout=min(left, 5.0)):
            if p.poll() is not None:
                break
            continue
        line = p.stdout.readline()
        if not line:
            break
        try:
            d = json.loads(line)
        except ValueError:
            continue
        if "step_s" in d:
            times.append(d["step_s"])
        if "done" in d:
            cores = d["cores"]
            break
    timed = times[warmup:] if len(times) > warmup else []
    if timed:
        return sum(timed) / len(timed), len(timed), warmup, cores
    if times:                                   # only warm-up steps finished inside the budget: report those (cold) steps
        return sum(times) / len(times), len(times), 0, cores
    return None, 0, 0, cores


def cpu_baseline(a, budget_s):
    dt, n, w, cores = run_cpu_bounded(a, 0, 1, budget_s)
    sample = ("B=1 utterance (T=%d,U=%d,V=%d) full train step incl. numpy front end, fp32, %d timed step(s), no separate warm-up, "
              "%d torch threads (capped: the fp32 CPU layers do not scale past a few dozen threads); oracle port of the reference "
              "path (torch-CPU layers + C lattice DP), warp_rnnt/PyKaldi being unavailable" % (a.T, a.U, a.V, n, cores))
    if dt is None:
        return {"value": None, "unit": "utt/s", "cores": cores, "kind": "port",
                "sample": sample + "; the step did not finish inside the %.0f s budget (value < %.4f utt/s)" % (budget_s, 1.0 / budget_s)}
    return {"value": 1.0 / dt, "unit": "utt/s", "cores": cores, "kind": "port", "sample": sample}


def run_reference(a):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    budget = 240.0
    dt, steps, warm, cores = run_cpu_bounded(a, min(a.warmup, 1), a.steps, budget)
    sample = ("B=1 utterance per step at the full (T=%d,U=%d,V=%d) shape; %d timed step(s) finished inside a %.0f s budget "
              "(requested %d), %d warm-up, %d torch threads" % (a.T, a.U, a.V, steps, budget, a.steps, warm, cores))
    val = (1.0 / dt) if dt else None
    print(json.dumps({"impl": "reference", "metric": METRIC, "value": val, "unit": "utt/s", "n_gpus": a.gpus, "steps": steps,
                      "warmup": warm, "ms_per_step": dt * 1e3 if dt else None, "higher_is_better": True, "scaling": "weak",
                      "vs_baseline": None, "dtype": "f32", "data": "synthetic",
                      "config": {"workload": "configs[1]: RNN-T train step batch=32/GPU T=%d U=%d V=%d; CPU arm runs B=1 samples" % (a.T, a.U, a.V)},
                      "cpu_baseline": {"value": val, "unit": "utt/s", "cores": cores, "kind": "port", "sample": sample},
                      "e2e": {"value": val, "unit": "utt/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}))


# ------------------------------------------------------------------------------------------------ GPU arm
def run_ours(a):
    import numpy as np
    import torch
    import torch.distributed as dist
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    assert torch.cuda.is_available(), "bench.py (ours) needs a GPU; there is no CPU fallback"
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group(backend="nccl", init_method="env://", device_id=dev)
    from pika_b200 import engine, _lib
    from pika_b200.frontend import FbankOptions, Frontend
    from pika_b200.model.transducer import Net
    from pika_b200.trainer.bmuf import BmufTrainer
    from pika_b200.trainer.flat import FlatParams, SgdNesterovClip
    from pika_b200.trainer.step import TrainStep
    from pika_b200.utils.spec_augment import SpecAugment

    engine.set_precision("bf16")
    engine.set_dropout_enabled(True)
    engine.set_seed(777 + rank)
    ta = train_args()
    torch.manual_seed(777)
    model = Net(model_args(a.V), 240, a.V).to(dev)
    model.train()
    flat = FlatParams(model)
    bmuf = BmufTrainer(0, rank, world, model, ta.block_momentum, ta.block_lr, flat=flat)
    opt = SgdNesterovClip(flat, ta.initial_lr, ta.momentum, ta.grad_clip)
    fe = Frontend(FbankOptions(num_mel_bins=80, low_freq=40.0, high_freq=-200.0, dither=0.0, window_type="hamming"), 1, 1, dev)
    torch.manual_seed(777 + rank)
    np.random.seed(777 + rank)
    step = TrainStep(model, ta, fe, bmuf, opt, spec_augmentor=SpecAugment(ta.max_freq_span, ta.max_time_span))

    B = a.batch
    pcm_host = torch.from_numpy(synth_pcm(B, a.T, 777 + rank)).pin_memory()
    rng = np.random.default_rng(777 + rank)
    tgt_host = torch.from_numpy(rng.integers(1, a.V, (B, a.U)).astype(np.int64)).pin_memory()
    n = pcm_host.shape[1]
    rate = [1.0] * B
    new_len, frames = Frontend.lengths([n] * B, rate)
    meta_host = torch.tensor([[n] * B, new_len, frames, [a.U] * B], dtype=torch.int32).pin_memory()
    fmeta_host = torch.tensor([rate, [-25.0] * B], dtype=torch.float32).pin_memory()
    t_max = max(frames)

    def to_device():
        pcm = pcm_host.to(dev, non_blocking=True)
        tgt = tgt_host.to(dev, non_blocking=True)
        meta = meta_host.to(dev, non_blocking=True)
        fmeta = fmeta_host.to(dev, non_blocking=True)
        return dict(pcm=pcm, target=tgt, n_samples=meta[0], new_len=meta[1], n_frames=meta[2], ali_lens=meta[3], rate=fmeta[0],
                    target_db=fmeta[1], t_max=t_max)
    h2d_bytes = pcm_host.numel() * 2 + tgt_host.numel() * 8 + meta_host.numel() * 4 + fmeta_host.numel() * 4
    resident = to_device()

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def timed(fn, steps):
        barrier()
        s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        s.record()
        for _ in range(steps):
            fn()
        e.record()
        barrier()
        ms = torch.tensor([s.elapsed_time(e)], device=dev)
        if world > 1:
            dist.all_reduce(ms, op=dist.ReduceOp.MAX)
        return ms.item()

    last = {}

    def step_device():
        last["c"] = step(resident)

    def step_e2e():
        costs = step(to_device())
        last["loss"] = float(costs.sum().item())                 # D2H read of the step's result

    for _ in range(max(a.warmup, 3)):
        step_device()
    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()
    l0 = _lib.launch_count()
    ms = timed(step_device, a.steps)
    launches = _lib.launch_count() - l0
    clocks = sampler.stop() if rank == 0 else None
    step_e2e()
    ms_e2e = timed(step_e2e, a.steps)
    value = world * B * a.steps / (ms / 1e3)
    e2e = world * B * a.steps / (ms_e2e / 1e3)

    if rank != 0:
        if world > 1:
            dist.barrier()
            dist.destroy_process_group()
        return
    hbm, tf_burst, tf_sus, src = peaks()
    # ---- live roofline of the dominant kernel: the fc2 joint GEMM (forward shape), timed alone with CUDA events
    Tp = int((frames[0] - 42 + 3) // 4)
    R, H = B * Tp * (a.U + 1), 1024
    from pika_b200 import kernels as K
    hh = (torch.randn(R, H, device=dev) * 0.3).to(torch.bfloat16)      # gate outputs tanh*sigmoid: |h| < 1
    w2 = (torch.randn(a.V, H, device=dev) * 0.03).to(torch.bfloat16)   # nn.Linear init scale 1/sqrt(1024)
    out = torch.empty(R, a.V, device=dev, dtype=torch.bfloat16)

    def ev_time(fn, it=5):
        fn(); torch.cuda.synchronize()
        s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        s.record()
        for _ in range(it):
            fn()
        e.record(); torch.cuda.synchronize()
        return s.elapsed_time(e) / it
    from pika_b200 import engine as _E
    fused_lse = bool(_E._FUSED_LSE) and a.V % 8 == 0
    b2 = torch.zeros(a.V, device=dev)
    parts = torch.empty((a.V + 255) // 256, R, 2, device=dev) if fused_lse else None
    g_ms_plain = ev_time(lambda: K.gemm(hh, w2, out, bias=b2, block_n=256))
    # the launch as the step issues it: with the fused row log-sum-exp partials when PK_FUSED_LSE is on
    g_ms = ev_time(lambda: K.gemm(hh, w2, out, bias=b2, block_n=256, row_lse=parts)) if fused_lse else g_ms_plain
    g_tf = 2.0 * R * H * a.V / g_ms / 1e9
    lab = torch.randint(1, a.V, (B, a.U), device=dev, dtype=torch.int32)
    fl = torch.full((B,), Tp, device=dev, dtype=torch.int32)
    ll = torch.full((B,), a.U, device=dev, dtype=torch.int32)
    z = out.view(B, Tp, a.U + 1, a.V)
    loss_passes = 2.0 if fused_lse else 3.0           # logits read for the gradient + dlogits write (+ the first-pass read when not fused)

    def loss_time(it=3):
        # the loss overwrites the logits in place, so every timed launch gets freshly produced logits (and partials)
        tot = 0.0
        for i in range(it + 1):
            K.gemm(hh, w2, out, bias=b2, block_n=256, row_lse=parts)
            s_, e_ = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            s_.record()
            K.rnnt_loss_fwd_bwd(z, lab, fl, ll, dlogits=z, row_lse=parts)
            e_.record(); torch.cuda.synchronize()
            if i > 0:
                tot += s_.elapsed_time(e_)
        return tot / it
    l_ms = loss_time()
    l_gbs = loss_passes * z.numel() * 2 / l_ms / 1e6
    del hh, out, z
    res = {
        "metric": METRIC, "value": value, "unit": "utt/s", "n_gpus": world, "steps": a.steps, "warmup": max(a.warmup, 3),
        "ms_per_step": ms / a.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "bf16",
        "data": "synthetic",
        "config": {"workload": "configs[1]: RNN-T train step (front end+encoder+pred+joint+loss+backward+clip/SGD, BMUF every 5th step) "
                               "batch=%d/GPU T=%d U=%d V=%d bf16; T'=%d" % (B, a.T, a.U, a.V, Tp),
                   "global_batch": world * B, "parallelism": "bmuf-dp%d" % world,
                   "l2": "no flush needed: each step streams >30 GB (13.9 GB logits alone) through a 126 MB L2",
                   "peaks": src},
        "clocks": clocks,
        "e2e": {"value": e2e, "unit": "utt/s", "h2d_bytes_per_step": h2d_bytes, "d2h_bytes_per_step": 4,
                "ms_per_step": ms_e2e / a.steps, "loss": last.get("loss")},
        "gpu_launches": launches,
        "roofline": {"kernel": "gemm_tcgen05_kernel (joint fc2 forward, M=%d N=%d K=%d, bias%s)" % (R, a.V, H, " + fused row-LSE epilogue" if fused_lse else ""), "bound": "tensor",
                     "achieved": g_tf, "peak": tf_burst, "unit": "TFLOP/s", "frac": g_tf / tf_burst,
                     "traffic": (NCU_FC2_LSE_TRAFFIC_BYTES if fused_lse else NCU_FC2_TRAFFIC_BYTES) if (R, a.V, H) == (1159680, 6000, 1024) else None,
                     "traffic_source": ("profiles/r01_gemm_fc2_lse.ncu.txt (dram read 4.67 GB + write 14.10 GB per launch" if fused_lse else
                                        "profiles/r01_gemm_fc2.ncu.txt (dram read 3.36 GB + write 13.88 GB per launch") +
                                       "; algorithmic: A 2.38 GB + B 0.012 GB + C 13.92 GB (+ 0.22 GB row partials when fused))",
                     "launch_ms": g_ms, "launch_ms_plain_epilogue": g_ms_plain},
        "roofline_loss": {"kernel": ("rnnt_rowfinish + rnnt_lattice + rnnt_grad (first pass done in the fc2 GEMM epilogue)" if fused_lse else
                                     "rnnt_rowstats + rnnt_lattice + rnnt_grad (fused log-softmax + RNN-T loss + gradient)"), "bound": "hbm",
                          "achieved": l_gbs, "peak": hbm, "unit": "GB/s", "frac": l_gbs / hbm, "traffic": None, "launch_ms": l_ms,
                          "algorithmic_bytes": loss_passes * B * Tp * (a.U + 1) * a.V * 2},
    }
    if world == 1 and not a.no_cpu_baseline:
        try:
            res["cpu_baseline"] = cpu_baseline(a, a.cpu_budget_s)
        except Exception as ex:                                   # the baseline is reported, never required
            res["cpu_baseline"] = {"error": repr(ex)}
    print(json.dumps(res))
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


if __name__ == "__main__":
    args = parse()
    if args.impl == "_cpu_worker":
        _cpu_worker(sys.argv[1:])
    elif args.impl == "reference":
        run_reference(args)
    else:
        run_ours(args)
