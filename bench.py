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
    ap.add_argument("--batch", type=int, default=None, help="per-GPU batch (default: 32 train, 64 decode, 16 mbr)")
    ap.add_argument("--T", type=int, default=None, help="fbank frames per utterance (default: 1000; 1500 for decode)")
    ap.add_argument("--U", type=int, default=150)
    ap.add_argument("--V", type=int, default=6000)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--cpu-budget-s", type=float, default=150.0)
    ap.add_argument("--cpu-threads", type=int, default=0, help="CPU arm: torch threads (0 = measure 32 and all cores, keep the faster)")
    ap.add_argument("--workload", default="train", choices=["train", "decode", "mbr"],
                    help="train = configs[1]/[2] (the headline metric); decode = configs[4] (beam 16, batch 64, T=1500; RTF); "
                         "mbr = configs[3] (MBR train step, batch 16 per GPU)")
    ap.add_argument("--beam", type=int, default=0)
    a = ap.parse_args()
    if a.batch is None:
        a.batch = {"train": 32, "decode": 64, "mbr": 16}[a.workload]
    if a.T is None:
        a.T = 1500 if a.workload == "decode" else 1000
    return a


def model_args(V):
    return types.SimpleNamespace(rnn_size=1024, local_rank=0, decoder_type="rnn", brnn=True, encoder_type="transformer",
                                 embd_dim=100, padding_idx=V, dropout=0.2, dec_layers=2, enc_layers=9)


def train_args():
    # egs/train_transducer_bmuf_otfaug.sh:157-197 (recipe values); SpecAugment on (north star)
    return types.SimpleNamespace(cmn=True, model_lctx=21, model_rctx=21, model_stride=4, sync_period=5, initial_lr=4e-4,
                                 final_lr=4e-5, momentum=0.9, grad_clip=3.0, num_epochs=15, num_batches_per_epoch=1000, epoch=0,
                                 block_momentum=0.9, block_lr=1.0, max_freq_span=15, max_time_span=35)


def tprime(T):
    return (T - 42 + 3) // 4


def train_workload(a):
    return ("configs[1]: RNN-T train step (front end+encoder+pred+joint+loss+backward+clip/SGD, BMUF every 5th step) "
            "batch=%d/GPU T=%d U=%d V=%d bf16; T'=%d" % (a.batch, a.T, a.U, a.V, tprime(a.T)))


def decode_workload(a):
    return ("configs[4]: batch beam-search decode beam=%d batch=%d T=%d (T'=%d) V=%d, 1 GPU" % (a.beam or 16, a.batch, a.T, tprime(a.T), a.V))


def mbr_workload(a):
    return ("configs[3]: MBR train step (N-best beam=%d decode + RNN-T branch + path-gathered MBR branch + clip/SGD, BMUF every 5th step) "
            "batch=%d/GPU T=%d U=%d V=%d bf16" % (a.beam or 4, a.batch, a.T, a.U, a.V))


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
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.gpu), "--query-gpu=" + self.Q, "--format=csv,noheader,nounits", "-lms", "100"],
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
        pw = sorted(float(r[3]) for r in self.rows if r[3].replace(".", "").isdigit())
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": max(mx) if mx else None, "reasons": sorted(reasons),
                "samples": len(self.rows), "sm_mhz_min": sm[0] if sm else None, "power_w": pw[len(pw) // 2] if pw else None}


# ------------------------------------------------------------------------------------------------ CPU arm
NCU_FC2_TRAFFIC_BYTES = 20.05e9        # dram__bytes_read.sum + dram__bytes_write.sum of one plain fc2 launch (profiles/r02_gemm_fc2_fwd.ncu.txt: 6.16 + 13.89 GB)
NCU_FC2_LSE_TRAFFIC_BYTES = 18.836e9   # same for the launch with the row-LSE epilogue (profiles/r02_gemm_fc2_fwd_lse.ncu.txt: 4.52 + 14.31 GB)
CPU_THREADS_CAP = 32      # first candidate thread count of the CPU arm; "all cores" is the second, the faster one is kept (measured per run)


def cpu_step_fn(a, seed=777):
    """Builds the oracle port of one training batch at B=1 (the bounded sample) and returns (fn, cores)."""
    import numpy as np
    import torch
    from oracle import train_step as ots
    from pika_b200.model.transducer import Net
    cores = a.cpu_threads if a.cpu_threads > 0 else min(os.cpu_count() or 1, CPU_THREADS_CAP)
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
        first = state["bufs"] is None
        costs, state["bufs"] = ots.train_step(sd, data, labels, tl.astype(np.int32), np.array([a.U], np.int32), ta.initial_lr,
                                              ta.momentum, ta.grad_clip, state["bufs"])
        if first:      # the first step starts from the reference initial weights: its loss and fc2-bias gradient (= column sums of
            # d loss / d logits over all T' x (U+1) rows) are what the GPU arm reproduces at B=1 (parity_full_shape)
            state["first"] = {"loss": float(costs.sum()), "db2": [float(v) for v in sd["fc2.bias"].grad.numpy()]}
        return float(costs.sum())
    fn.state = state
    return fn, cores


def _cpu_worker(argv):
    """subprocess entry: times `steps` CPU batches after `warm` warm-ups and prints one JSON line"""
    a = parse_from(argv)
    if a.workload == "decode":
        return _cpu_decode_worker(a)
    if a.workload == "mbr":
        return _cpu_mbr_worker(a)
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
            print(json.dumps({"first_step": fn.state["first"]}), flush=True)
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


def run_cpu_bounded(a, warmup, steps, budget_s, threads=0, extra=None):
    """Runs the CPU arm in a subprocess with a hard wall-clock budget; returns (mean step seconds over the timed
    steps that finished, number of timed steps, number of warm-ups, cores).  ``extra`` (dict) receives the worker's
    side-channel lines (first-step loss / gradient signature)."""
    cmd = [sys.executable, os.path.abspath(__file__), "--impl", "_cpu_worker", "--T", str(a.T), "--U", str(a.U), "--V", str(a.V),
           "--steps", str(steps), "--warmup", str(warmup), "--cpu-threads", str(threads), "--workload", a.workload,
           "--batch", str(a.batch), "--beam", str(a.beam)]
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
        if extra is not None and ("first_step" in d or "decode" in d or "mbr" in d):
            extra.update(d)
        if "done" in d:
            cores = d["cores"]
            break
    timed = times[warmup:] if len(times) > warmup else []
    if timed:
        return sum(timed) / len(timed), len(timed), warmup, cores
    if times:                                   # only warm-up steps finished inside the budget: report those (cold) steps
        return sum(times) / len(times), len(times), 0, cores
    return None, 0, 0, cores


def cpu_thread_candidates():
    allc = os.cpu_count() or 1
    return sorted({min(allc, CPU_THREADS_CAP), allc})


def cpu_baseline(a, budget_s, extra=None):
    """times the CPU arm at each candidate thread count (32 and all host cores) and keeps the faster: the choice is measured, per run"""
    cands = cpu_thread_candidates()
    tried, best = {}, None
    for th in cands:
        dt, n, w, cores = run_cpu_bounded(a, 1 if len(cands) == 1 else 0, 1, budget_s / len(cands), threads=th, extra=extra)
        tried[str(th)] = None if dt is None else round(1.0 / dt, 5)
        if dt is not None and (best is None or dt < best[0]):
            best = (dt, n, w, cores)
    sample = ("B=1 utterance (T=%d,U=%d,V=%d) full train step incl. numpy front end, fp32, one timed step per thread count, no separate "
              "warm-up; utt/s by torch thread count: %s (the faster is reported); oracle port of the reference path (torch-CPU layers + "
              "C lattice DP), warp_rnnt/PyKaldi being unavailable" % (a.T, a.U, a.V, json.dumps(tried)))
    if best is None:
        return {"value": None, "unit": "utt/s", "cores": cands[-1], "kind": "port",
                "sample": sample + "; no step finished inside the %.0f s budget" % budget_s}
    return {"value": 1.0 / best[0], "unit": "utt/s", "cores": best[3], "kind": "port", "sample": sample}


def run_reference(a):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    if a.workload == "decode":
        return run_reference_decode(a)
    if a.workload == "mbr":
        return run_reference_mbr(a)
    budget = 240.0
    # thread count: one probing step at each candidate, then the timed run with the faster
    probe = {}
    for th in cpu_thread_candidates():
        dtp, _, _, _ = run_cpu_bounded(a, 0, 1, 40.0, threads=th)
        probe[th] = dtp if dtp is not None else 1e9
    th_best = min(probe, key=probe.get)
    dt, steps, warm, cores = run_cpu_bounded(a, min(a.warmup, 1), a.steps, budget - 80.0, threads=th_best)
    sample = ("B=1 utterance per step at the full (T=%d,U=%d,V=%d) shape; %d timed step(s) finished inside a %.0f s budget "
              "(requested %d), %d warm-up, %d torch threads (probe step seconds by thread count: %s)"
              % (a.T, a.U, a.V, steps, budget - 80.0, a.steps, warm, cores, json.dumps({str(k): round(v, 2) for k, v in probe.items()})))
    val = (1.0 / dt) if dt else None
    print(json.dumps({"impl": "reference", "metric": METRIC, "value": val, "unit": "utt/s", "n_gpus": a.gpus, "steps": steps,
                      "warmup": warm, "ms_per_step": dt * 1e3 if dt else None, "higher_is_better": True, "scaling": "weak",
                      "vs_baseline": None, "dtype": "f32", "data": "synthetic",
                      "config": {"workload": train_workload(a), "global_batch": a.gpus * a.batch, "parallelism": "bmuf-dp%d" % a.gpus},
                      "cpu_baseline": {"value": val, "unit": "utt/s", "cores": cores, "kind": "port", "sample": sample},
                      "e2e": {"value": val, "unit": "utt/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}))


# ------------------------------------------------------------------------------------------------ GPU arm
def parity_full_shape(a, first, dev):
    """The utterance the CPU arm just pushed through the oracle (T, U, V of the workload, B=1, reference initial weights, fixed
    SpecAugment mask, dropout off), run through the GPU path in bf16: loss and the fc2 bias gradient (= the column sums of
    d loss / d logits over all T' x (U+1) joint rows, a V-wide signature of the fused loss gradient)."""
    import numpy as np
    import torch
    from pika_b200 import engine
    from pika_b200.frontend import FbankOptions, Frontend
    from pika_b200.model.transducer import Net
    seed = 777
    torch.manual_seed(seed)
    model = Net(model_args(a.V), 240, a.V).to(dev)
    model.train()
    fe = Frontend(FbankOptions(num_mel_bins=80, low_freq=40.0, high_freq=-200.0, dither=0.0, window_type="hamming"), 1, 1, dev)
    pcm = torch.from_numpy(synth_pcm(1, a.T, seed)).to(dev)
    labels = torch.from_numpy(np.random.default_rng(seed).integers(1, a.V, (1, a.U)).astype(np.int64)).to(dev)
    n = pcm.shape[1]
    new_len, frames = Frontend.lengths([n], [1.0])
    i32 = lambda v: torch.tensor(v, dtype=torch.int32, device=dev)                       # noqa: E731
    engine.set_dropout_enabled(False)
    try:
        feats = fe(pcm, i32([n]), torch.ones(1, device=dev), torch.full((1,), -25.0, device=dev), i32(new_len), i32(frames), max(frames),
                   out_dtype=engine.act_dtype(), cmn=True, offset=None, scale=None, specaug=(100, 7, 300, 20))
        tl = i32([tprime(frames[0])])
        costs = engine.transducer_loss(model, feats, labels, tl, i32([a.U]))
        costs.sum().backward()
    finally:
        engine.set_dropout_enabled(True)
    loss = float(costs.sum().item())
    db2 = model.fc2.bias.grad.double().cpu().numpy()
    ref = np.asarray(first["db2"], np.float64)
    return {"loss_rel": abs(loss - first["loss"]) / abs(first["loss"]), "dlogits_colsum_rel": float(np.linalg.norm(db2 - ref) / np.linalg.norm(ref)),
            "loss_gpu": loss, "loss_cpu_oracle": first["loss"],
            "what": "B=1 (T=%d,U=%d,V=%d) utterance: GPU bf16 path vs the CPU oracle port (fp32), same PCM, labels, initial weights and "
                    "SpecAugment mask, dropout off" % (a.T, a.U, a.V)}


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
    # the recipe always passes --cmvn_stats (egs/train_transducer_bmuf_otfaug.sh): CMN + global offset/scale; identity statistics here
    cmvn_off, cmvn_scale = torch.zeros(fe.D, device=dev), torch.ones(fe.D, device=dev)
    step = TrainStep(model, ta, fe, bmuf, opt, offset=cmvn_off, scale=cmvn_scale, spec_augmentor=SpecAugment(ta.max_freq_span, ta.max_time_span))

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

    host = {}

    def timed(fn, steps):
        barrier()
        s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        s.record()
        t0 = time.perf_counter()
        for _ in range(steps):
            fn()
        host["issue_ms"] = (time.perf_counter() - t0) * 1e3 / steps     # host time to ISSUE a step (no sync inside the loop)
        e.record()
        barrier()
        ms = torch.tensor([s.elapsed_time(e)], device=dev)
        if world > 1:
            allms = [torch.zeros_like(ms) for _ in range(world)]
            dist.all_gather(allms, ms)
            host["per_rank_ms"] = [round(float(t.item()) / steps, 3) for t in allms]
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
    host_issue_ms = host["issue_ms"]
    per_rank_ms = host.get("per_rank_ms")
    one = []
    for _ in range(3):              # pure host cost of a step: issued into an EMPTY launch queue (a step is < 1024 launches), then drained
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        step_device()
        one.append((time.perf_counter() - t0) * 1e3)
        torch.cuda.synchronize()
    host_step_ms = sorted(one)[1]
    clocks = sampler.stop() if rank == 0 else None
    step_e2e()
    ms_e2e = timed(step_e2e, a.steps)
    value = world * B * a.steps / (ms / 1e3)
    e2e = world * B * a.steps / (ms_e2e / 1e3)
    # duration of the two roofline kernels INSIDE real steps: CUDA events around those launches on the launching stream (engine.EVENT_TAPS)
    from pika_b200 import engine as _Eng
    _Eng.EVENT_TAPS = {}
    for _ in range(min(a.steps, 5)):
        step_device()
    torch.cuda.synchronize()
    taps = {k: sum(s_.elapsed_time(e_) for s_, e_ in v) / len(v) for k, v in _Eng.EVENT_TAPS.items() if v}
    _Eng.EVENT_TAPS = None

    # multi-GPU correctness, not just speed: after one more block sync every rank must hold bit-identical parameters
    params_identical = None
    sync_ms = None
    if world > 1:
        sync_ms = []
        for _ in range(3):                        # one BMUF block sync (delta, all-reduce of the flat parameter vector, update, NaN vote) timed alone
            barrier()
            s_, e_ = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            s_.record(); bmuf.update_and_sync(); e_.record()
            torch.cuda.synchronize()
            sync_ms.append(round(s_.elapsed_time(e_), 3))
        bmuf.update_and_sync()
        sig = torch.stack([flat.data.double().sum(), flat.data.double().abs().sum(), flat.data[::9973].double().square().sum()])
        sigs = [torch.zeros_like(sig) for _ in range(world)]
        dist.all_gather(sigs, sig)
        params_identical = bool(all(torch.equal(x_, sigs[0]) for x_ in sigs))

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
    parts = torch.empty(K.row_lse_parts(R, a.V, 256), R, 2, device=dev) if fused_lse else None
    g_ms_plain = ev_time(lambda: K.gemm(hh, w2, out, bias=b2, block_n=256))
    # the launch as the step issues it: with the fused row log-sum-exp partials when PK_FUSED_LSE is on
    g_ms = ev_time(lambda: K.gemm(hh, w2, out, bias=b2, block_n=256, row_lse=parts)) if fused_lse else g_ms_plain
    g_ms_alone = g_ms
    if "fc2_fwd" in taps:
        g_ms = taps["fc2_fwd"]                        # the launch inside the step (the stand-alone loop above runs hotter: reported beside it)
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
    l_ms_alone = loss_time()
    l_ms = taps.get("rnnt_loss", l_ms_alone)          # rowfinish + lattice + grad (+ column-sum partials) inside the step
    l_gbs = loss_passes * z.numel() * 2 / l_ms / 1e6
    del hh, out, z
    res = {
        "metric": METRIC, "value": value, "unit": "utt/s", "n_gpus": world, "steps": a.steps, "warmup": max(a.warmup, 3),
        "ms_per_step": ms / a.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "bf16",
        "data": "synthetic",
        "config": {"workload": train_workload(a), "global_batch": world * B, "parallelism": "bmuf-dp%d" % world,
                   "l2": "no flush needed: each step streams >30 GB (13.9 GB logits alone) through a 126 MB L2",
                   "peaks": src},
        "clocks": clocks,
        "e2e": {"value": e2e, "unit": "utt/s", "h2d_bytes_per_step": h2d_bytes, "d2h_bytes_per_step": 4,
                "ms_per_step": ms_e2e / a.steps, "loss": last.get("loss")},
        "gpu_launches": launches, "host_ms_per_step": {"python_issue_empty_queue": round(host_step_ms, 3), "in_timed_loop": round(host_issue_ms, 3)},
        "per_rank_ms_per_step": per_rank_ms, "bmuf_sync_ms": sync_ms,
        "roofline": {"kernel": "gemm_tcgen05_kernel (joint fc2 forward, M=%d N=%d K=%d, bias%s)" % (R, a.V, H, " + fused row-LSE epilogue" if fused_lse else ""), "bound": "tensor",
                     "achieved": g_tf, "peak": (tf_sus if "fc2_fwd" in taps else tf_burst), "unit": "TFLOP/s",
                     "frac": g_tf / (tf_sus if "fc2_fwd" in taps else tf_burst),
                     "peak_kind": ("MEASURED_PEAKS sustained cuBLAS bf16 figure: the kernel is timed inside the long step" if "fc2_fwd" in taps
                                   else "MEASURED_PEAKS burst cuBLAS bf16 figure: the kernel is timed alone"),
                     "alone": {"launch_ms": g_ms_alone, "achieved": 2.0 * R * H * a.V / g_ms_alone / 1e9, "peak": tf_burst,
                               "frac": 2.0 * R * H * a.V / g_ms_alone / 1e9 / tf_burst, "what": "same launch in a loop of its own, against the burst figure"},
                     "traffic": (NCU_FC2_LSE_TRAFFIC_BYTES if fused_lse else NCU_FC2_TRAFFIC_BYTES) if (R, a.V, H) == (1159680, 6000, 1024) else None,
                     "traffic_source": ("profiles/r02_gemm_fc2_fwd_lse.ncu.txt (one `ncu --set full` capture of this kernel at this shape: dram read 4.52 GB + write 14.31 GB per launch" if fused_lse else
                                        "profiles/r02_gemm_fc2_fwd.ncu.txt (one `ncu --set full` capture: dram read 6.16 GB + write 13.89 GB per launch") +
                                       "; a committed constant, not measured in this run; algorithmic: A 2.38 GB + B 0.012 GB + C 13.92 GB (+ 0.45 GB row partials when fused))",
                     "launch_ms": g_ms, "launch_ms_source": ("CUDA events around this launch inside %d real steps" % min(a.steps, 5)) if "fc2_fwd" in taps else "stand-alone loop",
                     "launch_ms_plain_epilogue_alone": g_ms_plain},
        "roofline_loss": {"kernel": ("rnnt_rowfinish + rnnt_lattice + rnnt_grad (first pass done in the fc2 GEMM epilogue)" if fused_lse else
                                     "rnnt_rowstats + rnnt_lattice + rnnt_grad (fused log-softmax + RNN-T loss + gradient)"), "bound": "hbm",
                          "achieved": l_gbs, "peak": hbm, "unit": "GB/s", "frac": l_gbs / hbm, "traffic": None, "launch_ms": l_ms, "launch_ms_alone_in_a_loop": l_ms_alone,
                          "algorithmic_bytes": loss_passes * B * Tp * (a.U + 1) * a.V * 2},
    }
    if params_identical is not None:
        res["params_identical_after_sync"] = params_identical
    if world == 1 and not a.no_cpu_baseline:
        try:
            extra = {}
            res["cpu_baseline"] = cpu_baseline(a, a.cpu_budget_s, extra)
            if "first_step" in extra:
                res["parity_full_shape"] = parity_full_shape(a, extra["first_step"], dev)
        except Exception as ex:                                   # the baseline is reported, never required
            res["cpu_baseline"] = {"error": repr(ex)}
    print(json.dumps(res))
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()



# ------------------------------------------------------------------------------------------------ decode workload (configs[4])
DECODE_METRIC = "decode RTF (batch beam search, beam=16, batch=64, T=1500, V=6000)"


def decode_model(a, dev):
    import torch
    from pika_b200.model.transducer import Net
    torch.manual_seed(777)
    m = Net(model_args(a.V), 240, a.V)
    with torch.no_grad():
        m.fc2.bias[0] += 6.0          # a randomly initialised transducer never emits blank: bias it so that alignments consume frames
    return m.to(dev).eval() if dev is not None else m.eval()


def decode_feats(B, T, seed):
    import numpy as np
    return np.random.default_rng(seed).standard_normal((B, T, 240)).astype(np.float32)


def _cpu_decode_worker(a):
    """CPU arm of the decode workload: the oracle port of decode_batch (oracle/decode.py, pinned bit-exactly to the reference's own
    outputs) on a bounded sample of the same workload -- 2 utterances of the same length, same beam -- encoder included."""
    import torch
    from oracle import decode as od, model as om
    cores = a.cpu_threads if a.cpu_threads > 0 else (os.cpu_count() or 1)
    torch.set_num_threads(cores)
    m = decode_model(a, None)
    sd = {k: v.detach() for k, v in m.state_dict().items()}
    nb = 2
    x = torch.from_numpy(decode_feats(nb, a.T, 1))
    tl = [tprime(a.T)] * nb
    t0 = time.time()
    with torch.no_grad():
        enc = om.encoder_forward(sd, x, train=False)
    ret = od.decode_batch(sd, enc, tl, a.beam or 16, n_best=1, max_len=[t + 100 for t in tl])
    dt = time.time() - t0
    audio = nb * (400 + (a.T - 1) * 160) / 16000.0
    print(json.dumps({"decode": {"rtf": dt / audio, "wall_s": dt, "audio_s": audio, "utts": nb, "cores": cores,
                                 "steps": max(len(h[0]) for h in ret["predictions"]) + 1}}), flush=True)
    print(json.dumps({"progress": 1, "step_s": dt}), flush=True)
    print(json.dumps({"done": 1, "first_s": dt, "total_s": dt, "cores": cores}), flush=True)


def cpu_baseline_decode(a, budget_s):
    """the beam loop is a chain of small products (rows = 2 x beam): torch's intra-op pool only hurts beyond a few threads (all 128
    host threads did not finish 2 utterances in 150 s on the GPU box), so the thread count is measured: 8 and 32, the faster is reported"""
    allc = os.cpu_count() or 1
    cands = sorted({min(allc, 8), min(allc, CPU_THREADS_CAP)})
    tried, best = {}, None
    for th in cands:
        extra = {}
        run_cpu_bounded(a, 0, 1, budget_s / len(cands), threads=th, extra=extra)
        d = extra.get("decode")
        tried[str(th)] = None if d is None else round(d["rtf"], 5)
        if d is not None and (best is None or d["rtf"] < best["rtf"]):
            best = d
    sample = ("2 utterances of the workload (T=%d, beam=%d, V=%d) through the oracle port of decode_batch (torch-CPU fp32 encoder / "
              "prediction net / joint, Python beam bookkeeping as in the reference); RTF by torch thread count: %s (the faster is reported)"
              % (a.T, a.beam or 16, a.V, json.dumps(tried)))
    if best is None:
        return {"value": None, "unit": "RTF", "cores": cands[-1], "kind": "port", "sample": sample + "; did not finish in %.0f s" % budget_s}
    return {"value": best["rtf"], "unit": "RTF", "cores": best["cores"], "kind": "port",
            "sample": sample + "; %d beam steps in %.1f s" % (best["steps"], best["wall_s"])}


def _cpu_mbr_worker(a):
    """CPU arm of the MBR workload: one utterance (the bounded sample) through the oracle ports of its three parts -- N-best generation
    (oracle/decode.py: beam, n_best = beam, no pruning, eval-mode encoder), the shared-encoder RNN-T branch and the path-gathered MBR
    branch with backward (oracle/mbr.py, pinned to the reference's loop body), inf-norm clip + Nesterov SGD (oracle/train.py)."""
    import numpy as np
    import torch
    from oracle import decode as od, mbr as ombr, model as om, train as ot
    cores = a.cpu_threads if a.cpu_threads > 0 else min(os.cpu_count() or 1, CPU_THREADS_CAP)
    torch.set_num_threads(cores)
    beam = a.beam or 4
    m = decode_model(a, None)
    sd = {k: v.detach().clone().requires_grad_(v.dtype.is_floating_point and "running" not in k) for k, v in m.state_dict().items()}
    x = torch.from_numpy(decode_feats(1, a.T, 11))
    tgt = torch.from_numpy(np.random.default_rng(3).integers(1, a.V, (1, a.U)).astype(np.int64))
    tl, ul = np.array([tprime(a.T)], np.int32), np.array([a.U], np.int32)
    t0 = time.time()
    with torch.no_grad():
        enc = om.encoder_forward({k: v.detach() for k, v in sd.items()}, x, train=False)
        ret = od.decode_batch({k: v.detach() for k, v in sd.items()}, enc, [int(tl[0])], beam, n_best=beam, max_len=[int(tl[0]) + a.U + 3],
                              sm_scale=0.8, beam_prune=False)
    t_dec = time.time() - t0
    hyps = [[h + [-1] for h in ret["predictions"][0]]]                     # the trainer sees alignments that end in EOS (:176-181)
    hyps = [[[t for t in h if t != -1] for h in hyps[0]]]
    mbr, costs = ombr.mbr_loss_and_grads(sd, x, tgt, tl, ul, hyps, [ret["scores"][0]], 0, a.V, 0.5, 0.8)
    params = [v for v in sd.values() if v.requires_grad and v.grad is not None]
    tot, coef = ot.clip_coef_inf([p.grad.numpy() for p in params], 3.0)
    with torch.no_grad():
        for p_ in params:
            p_.add_(p_.grad * float(coef), alpha=-1e-5)                    # first SGD step: the momentum buffer equals the gradient
    dt = time.time() - t0
    print(json.dumps({"mbr": {"utt_s": 1.0 / dt, "wall_s": dt, "decode_s": t_dec, "cores": cores, "mbr_loss": mbr,
                              "hyp_len": [len(h) for h in hyps[0]]}}), flush=True)
    print(json.dumps({"progress": 1, "step_s": dt}), flush=True)
    print(json.dumps({"done": 1, "first_s": dt, "total_s": dt, "cores": cores}), flush=True)


def cpu_baseline_mbr(a, budget_s):
    extra = {}
    run_cpu_bounded(a, 0, 1, budget_s, threads=min(os.cpu_count() or 1, CPU_THREADS_CAP), extra=extra)
    d = extra.get("mbr")
    sample = ("B=1 utterance (T=%d,U=%d,V=%d, beam %d) through the oracle ports of the MBR batch: N-best generation, RNN-T branch, MBR branch, "
              "backward, clip + SGD (torch-CPU fp32, %d threads), one step, no warm-up" % (a.T, a.U, a.V, a.beam or 4, min(os.cpu_count() or 1, CPU_THREADS_CAP)))
    if d is None:
        return {"value": None, "unit": "utt/s", "cores": min(os.cpu_count() or 1, CPU_THREADS_CAP), "kind": "port",
                "sample": sample + "; did not finish in %.0f s" % budget_s}
    return {"value": d["utt_s"], "unit": "utt/s", "cores": d["cores"], "kind": "port",
            "sample": sample + "; %.1f s of which N-best generation %.1f s" % (d["wall_s"], d["decode_s"])}


def run_reference_mbr(a):
    cb = cpu_baseline_mbr(a, 280.0)
    print(json.dumps({"impl": "reference", "metric": MBR_METRIC, "value": cb["value"], "unit": "utt/s", "n_gpus": a.gpus, "steps": 1, "warmup": 0,
                      "ms_per_step": (1e3 / cb["value"]) if cb["value"] else None, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
                      "dtype": "f32", "data": "synthetic", "config": {"workload": mbr_workload(a)}, "cpu_baseline": cb,
                      "e2e": {"value": cb["value"], "unit": "utt/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}))


def run_reference_decode(a):
    cb = cpu_baseline_decode(a, 280.0)
    print(json.dumps({"impl": "reference", "metric": DECODE_METRIC, "value": cb["value"], "unit": "RTF", "n_gpus": a.gpus, "steps": 1, "warmup": 0,
                      "higher_is_better": False, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
                      "config": {"workload": decode_workload(a)}, "cpu_baseline": cb,
                      "e2e": {"value": cb["value"], "unit": "RTF", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}))


def run_decode(a):
    """configs[4]: one step = decode_batch over one batch of 64 utterances x 1500 frames (encoder + beam search + back-trace).
    Replicas only across GPUs (utterance batches are independent): every rank decodes its own batch."""
    import numpy as np
    import torch
    import torch.distributed as dist
    rank, local_rank, world = int(os.environ.get("RANK", "0")), int(os.environ.get("LOCAL_RANK", "0")), int(os.environ.get("WORLD_SIZE", "1"))
    assert torch.cuda.is_available(), "bench.py needs a GPU; there is no CPU fallback"
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group(backend="nccl", init_method="env://", device_id=dev)
    from pika_b200 import engine, _lib
    from pika_b200.decoder.beam_transducer import GlobalScorer
    from pika_b200.decoder.transducer_decoder import TransducerDecoder
    engine.set_precision("bf16")
    beam, B = a.beam or 16, a.batch
    model = decode_model(a, dev)
    dargs = types.SimpleNamespace(las_rescorer=None, las_rescorer_bw=None, bilas_rescorer=None, nonblk_reward=0.0)
    dec = TransducerDecoder(model, B, beam, n_best=1, blk=0, global_scorer=GlobalScorer(), cuda=True, beam_prune=True, args=dargs)
    x_host = torch.from_numpy(decode_feats(B, a.T, 1 + rank)).pin_memory()
    x_res = x_host.to(dev)
    tl = torch.full((B,), tprime(a.T), dtype=torch.int32)
    ml = [int(t) + 100 for t in tl]
    audio_s = B * (400 + (a.T - 1) * 160) / 16000.0
    last = {}

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def timed(fn, steps):
        barrier()
        s_, e_ = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        s_.record()
        for _ in range(steps):
            fn()
        e_.record()
        barrier()
        ms = torch.tensor([s_.elapsed_time(e_)], device=dev)
        if world > 1:
            dist.all_reduce(ms, op=dist.ReduceOp.MAX)
        return ms.item()

    replayed = {"kernels": 0, "graphs": 0}

    def count_replays():
        replayed["kernels"] += getattr(dec, "last_replays", 0) * getattr(dec, "kernels_per_replay", 0)
        replayed["graphs"] += getattr(dec, "last_replays", 0)

    def step_device():
        last["ret"], _ = dec.decode_batch(x_res, tl, ml)
        count_replays()

    def step_e2e():
        last["ret"], _ = dec.decode_batch(x_host.to(dev, non_blocking=True), tl, ml)      # H2D of the features; the hypotheses come back inside

    for _ in range(max(a.warmup, 3)):
        step_device()
    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()
    l0 = _lib.launch_count()
    replayed["kernels"] = replayed["graphs"] = 0
    ms = timed(step_device, a.steps)
    host_launches = _lib.launch_count() - l0                    # launches issued by the host (eager steps + graph captures)
    launches = host_launches + replayed["kernels"]              # + kernels executed from replayed CUDA graphs
    graph_replays = replayed["graphs"]
    clocks = sampler.stop() if rank == 0 else None
    ms_e2e = timed(step_e2e, a.steps)
    if rank != 0:
        if world > 1:
            dist.barrier(); dist.destroy_process_group()
        return
    ret = last["ret"]
    beam_steps = max(len(h[0]) for h in ret["predictions"]) + 1
    rtf = (ms / 1e3 / a.steps) / audio_s
    rtf_e2e = (ms_e2e / 1e3 / a.steps) / audio_s
    hbm, _, _, src = peaks()
    rows, H, V = B * beam, 1024, a.V
    # algorithmic HBM bytes of one beam step: every weight matrix of the prediction net and the joint is read once (bf16), the
    # [rows, V] logits are written and read as f32, log-probs written and read as f32, states / activations are noise next to that
    w_bytes = 2 * (4 * H * (104 + H) + 4 * H * 2 * H + 2 * H * 2 * H + V * H)
    step_bytes = w_bytes + 4 * rows * V * 4
    step_ms = ms / a.steps / beam_steps
    res = {"metric": DECODE_METRIC, "value": rtf, "unit": "RTF", "n_gpus": world, "steps": a.steps, "warmup": max(a.warmup, 3),
           "ms_per_step": ms / a.steps, "higher_is_better": False, "scaling": "weak", "vs_baseline": None, "dtype": "bf16", "data": "synthetic",
           "config": {"workload": decode_workload(a), "parallelism": "replicas x%d" % world, "beam_steps": beam_steps,
                      "l2": "inputs are device resident; the per-step working set (weights 45 MB + logits) is meant to stay in the 126 MB L2", "peaks": src},
           "clocks": clocks,
           "e2e": {"value": rtf_e2e, "unit": "RTF", "h2d_bytes_per_step": x_host.numel() * 4,
                   "d2h_bytes_per_step": int((beam_steps * 2 + 3) * B * beam * 4), "ms_per_step": ms_e2e / a.steps},
           "gpu_launches": launches,
           "roofline": {"kernel": "one beam step (gather, 2-layer LSTM step, factored joint, fc2, log-softmax, beam advance, state reorder)",
                        "bound": "hbm", "achieved": step_bytes / step_ms / 1e6, "peak": hbm, "unit": "GB/s",
                        "frac": step_bytes / step_ms / 1e6 / hbm, "traffic": None, "launch_ms": step_ms, "algorithmic_bytes": step_bytes,
                        "note": "latency bound: %d kernels per beam step, %.2f host-side launches (graph replays + eager) per beam step"
                                % (launches // a.steps // max(beam_steps, 1), (graph_replays + host_launches - graph_replays * 0) / a.steps / max(beam_steps, 1))},
           "host_launches_per_beam_step": (graph_replays + host_launches) / a.steps / max(beam_steps, 1)}
    if world == 1 and not a.no_cpu_baseline:
        try:
            res["cpu_baseline"] = cpu_baseline_decode(a, a.cpu_budget_s)
        except Exception as ex:
            res["cpu_baseline"] = {"error": repr(ex)}
    print(json.dumps(res))
    if world > 1:
        dist.barrier(); dist.destroy_process_group()


# ------------------------------------------------------------------------------------------------ MBR workload (configs[3])
MBR_METRIC = "utterances/sec MBR train step (N-best decode + RNN-T loss + MBR loss, T=1000,U=150,V=6k)"


def run_mbr(a):
    """configs[3]: one step = N-best generation (beam 4, n_best = beam, no pruning) + shared-encoder RNN-T branch + path-gathered MBR
    branch + backward + clip/SGD; BMUF block sync every 5th step over NCCL.  Features are given (the MBR recipe shares the front end
    with configs[1], which bench.py --workload train measures)."""
    import numpy as np
    import torch
    import torch.distributed as dist
    rank, local_rank, world = int(os.environ.get("RANK", "0")), int(os.environ.get("LOCAL_RANK", "0")), int(os.environ.get("WORLD_SIZE", "1"))
    assert torch.cuda.is_available(), "bench.py needs a GPU; there is no CPU fallback"
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group(backend="nccl", init_method="env://", device_id=dev)
    from pika_b200 import engine, _lib
    from pika_b200.decoder.beam_transducer import GlobalScorer
    from pika_b200.decoder.transducer_decoder import TransducerDecoder
    from pika_b200.trainer.bmuf import BmufTrainer
    from pika_b200.trainer.flat import FlatParams, SgdNesterovClip
    from pika_b200.trainer.mbr import mbr_forward_backward
    engine.set_precision("bf16")
    engine.set_seed(777 + rank)
    beam, B = a.beam or 4, a.batch
    model = decode_model(a, dev)
    flat = FlatParams(model)
    bmuf = BmufTrainer(0, rank, world, model, 0.9, 1.0, flat=flat)
    opt = SgdNesterovClip(flat, 1e-5, 0.9, 3.0)
    dargs = types.SimpleNamespace(las_rescorer=None, las_rescorer_bw=None, bilas_rescorer=None, nonblk_reward=0.0)
    dec = TransducerDecoder(model, B, beam, n_best=beam, blk=0, global_scorer=GlobalScorer(), cuda=True, beam_prune=False, args=dargs)
    x_host = torch.from_numpy(decode_feats(B, a.T, 11 + rank)).pin_memory()
    tgt_host = torch.from_numpy(np.random.default_rng(3 + rank).integers(1, a.V, (B, a.U)).astype(np.int64)).pin_memory()
    x_res, tgt_res = x_host.to(dev), tgt_host.to(dev)
    tl = torch.full((B,), tprime(a.T), dtype=torch.int32)
    ul = torch.full((B,), a.U, dtype=torch.int32)
    tl_d, ul_d = tl.to(dev), ul.to(dev)
    ml = [int(t) + int(u) + 3 for t, u in zip(tl, ul)]
    state = {"n": 0}

    def step(x, tgt):
        model.eval()
        ret, _ = dec.decode_batch(x, tl, ml)
        model.train()
        flat.zero_grad()
        mbr, costs = mbr_forward_backward(model, x, tgt, tl_d, ul_d, ret, blk=0, rnnt_scale=0.5, sm_scale=0.8)
        opt.step()
        if state["n"] != 0 and state["n"] % 5 == 0:
            bmuf.update_and_sync()
            opt.reset()
        state["n"] += 1
        return mbr

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def timed(fn, steps):
        barrier()
        s_, e_ = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        s_.record()
        for _ in range(steps):
            fn()
        e_.record()
        barrier()
        ms = torch.tensor([s_.elapsed_time(e_)], device=dev)
        if world > 1:
            dist.all_reduce(ms, op=dist.ReduceOp.MAX)
        return ms.item()
    last = {}

    def step_device():
        last["mbr"] = step(x_res, tgt_res)

    def step_e2e():
        last["mbr"] = float(step(x_host.to(dev, non_blocking=True), tgt_host.to(dev, non_blocking=True)))     # the MBR loss is a host float: D2H inside

    for _ in range(max(a.warmup, 3)):
        step_device()
    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()
    l0 = _lib.launch_count()
    ms = timed(step_device, a.steps)
    launches = _lib.launch_count() - l0
    clocks = sampler.stop() if rank == 0 else None
    ms_e2e = timed(step_e2e, a.steps)
    params_identical = None
    if world > 1:
        bmuf.update_and_sync()
        sig = torch.stack([flat.data.double().sum(), flat.data.double().abs().sum()])
        sigs = [torch.zeros_like(sig) for _ in range(world)]
        dist.all_gather(sigs, sig)
        params_identical = bool(all(torch.equal(x_, sigs[0]) for x_ in sigs))
    if rank != 0:
        if world > 1:
            dist.barrier(); dist.destroy_process_group()
        return
    _, tf_burst, tf_sus, src = peaks()
    res = {"metric": MBR_METRIC, "value": world * B * a.steps / (ms / 1e3), "unit": "utt/s", "n_gpus": world, "steps": a.steps,
           "warmup": max(a.warmup, 3), "ms_per_step": ms / a.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
           "dtype": "bf16", "data": "synthetic",
           "config": {"workload": mbr_workload(a), "global_batch": world * B, "parallelism": "bmuf-dp%d" % world,
                      "l2": "no flush needed: each step streams several GB of joint logits through the 126 MB L2", "peaks": src},
           "clocks": clocks,
           "e2e": {"value": world * B * a.steps / (ms_e2e / 1e3), "unit": "utt/s", "h2d_bytes_per_step": x_host.numel() * 4 + tgt_host.numel() * 8,
                   "d2h_bytes_per_step": 4, "ms_per_step": ms_e2e / a.steps, "mbr_loss": last.get("mbr")},
           "gpu_launches": launches}
    if params_identical is not None:
        res["params_identical_after_sync"] = params_identical
    if world == 1 and not a.no_cpu_baseline:
        try:
            res["cpu_baseline"] = cpu_baseline_mbr(a, a.cpu_budget_s)
        except Exception as ex:
            res["cpu_baseline"] = {"error": repr(ex)}
    print(json.dumps(res))
    if world > 1:
        dist.barrier(); dist.destroy_process_group()


if __name__ == "__main__":
    args = parse()
    if args.impl == "_cpu_worker":
        _cpu_worker(sys.argv[1:])
    elif args.impl == "reference":
        run_reference(args)
    elif args.workload == "decode":
        run_decode(args)
    elif args.workload == "mbr":
        run_mbr(args)
    else:
        run_ours(args)
