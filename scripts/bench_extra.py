"""Secondary measurements for BASELINE configs 4 (MBR train step) and 5 (batch beam-search decode, RTF).
bench.py stays the contract for the headline metric (config 2); this script prints one JSON line per config."""
import sys, os, json, time, types
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import bench
from pika_b200 import engine, _lib
from pika_b200.model.transducer import Net
from pika_b200.decoder.transducer_decoder import TransducerDecoder
from pika_b200.decoder.beam_transducer import GlobalScorer
from pika_b200.trainer.flat import FlatParams, SgdNesterovClip
from pika_b200.trainer.mbr import mbr_forward_backward

dev = torch.device("cuda", 0)
V = 6000
which = sys.argv[1:] or ["decode", "mbr"]
engine.set_precision("bf16")
torch.manual_seed(777)
model = Net(bench.model_args(V), 240, V).to(dev)
with torch.no_grad():
    model.fc2.bias[0] += 6.0          # random init never emits blank: bias it so that alignments consume frames
dargs = types.SimpleNamespace(las_rescorer=None, las_rescorer_bw=None, bilas_rescorer=None, nonblk_reward=0.0)


def feats(B, T, seed):
    g = torch.Generator().manual_seed(seed)
    return torch.randn(B, T, 240, generator=g).to(dev)


def tprime(T):
    return (T - 42 + 3) // 4


if "decode" in which:
    B, T, beam = 64, 1500, 16
    model.eval()
    x = feats(B, T, 1)
    tl = torch.full((B,), tprime(T), dtype=torch.int32)
    dec = TransducerDecoder(model, B, beam, n_best=1, blk=0, global_scorer=GlobalScorer(), cuda=True, beam_prune=True, args=dargs)
    ml = [int(t) + 100 for t in tl]
    dec.decode_batch(x, tl, ml)       # warm-up
    torch.cuda.synchronize()
    l0 = _lib.launch_count()
    t0 = time.time()
    ret, _ = dec.decode_batch(x, tl, ml)
    torch.cuda.synchronize()
    dt = time.time() - t0
    audio_s = B * (400 + (T - 1) * 160) / 16000.0
    steps = max(len(h[0]) for h in ret["predictions"]) + 1
    print(json.dumps({"metric": "decode RTF (batch beam search, beam=16, batch=64, T=1500, V=6000)", "value": dt / audio_s, "unit": "RTF",
                      "higher_is_better": False, "n_gpus": 1, "wall_s": dt, "audio_s": audio_s, "beam_steps": steps, "dtype": "bf16",
                      "gpu_launches": _lib.launch_count() - l0, "data": "synthetic",
                      "config": {"workload": "configs[4]: batch beam-search decode beam=16 batch=64 T=1500 (T'=365) V=6000"}}), flush=True)

if "mbr" in which:
    B, T, U, beam = 16, 1000, 150, 4
    flat = FlatParams(model)
    opt = SgdNesterovClip(flat, 1e-5, 0.9, 3.0)
    x = feats(B, T, 2)
    tl = torch.full((B,), tprime(T), dtype=torch.int32)
    ul = torch.full((B,), U, dtype=torch.int32)
    tgt = torch.randint(1, V, (B, U), generator=torch.Generator().manual_seed(3))
    dec = TransducerDecoder(model, B, beam, n_best=beam, blk=0, global_scorer=GlobalScorer(), cuda=True, beam_prune=False, args=dargs)

    def step():
        model.eval()
        ret, _ = dec.decode_batch(x, tl, [int(t) + int(u) + 3 for t, u in zip(tl, ul)])
        model.train()
        flat.zero_grad()
        mbr, costs = mbr_forward_backward(model, x, tgt.to(dev), tl.to(dev), ul.to(dev), ret, blk=0, rnnt_scale=0.5, sm_scale=0.8)
        opt.step()
        return mbr
    step(); torch.cuda.synchronize()
    l0 = _lib.launch_count()
    t0 = time.time()
    n = 3
    for _ in range(n):
        mbr = step()
    torch.cuda.synchronize()
    dt = (time.time() - t0) / n
    print(json.dumps({"metric": "utterances/sec MBR train step (N-best decode + RNN-T branch + path-gathered MBR branch)", "value": B / dt,
                      "unit": "utt/s", "higher_is_better": True, "n_gpus": 1, "ms_per_step": dt * 1e3, "dtype": "bf16", "data": "synthetic",
                      "gpu_launches": (_lib.launch_count() - l0) // n, "mbr_loss": mbr,
                      "config": {"workload": "configs[3] per-GPU shape: MBR train step batch=16 T=1000 U=150 V=6000 beam=4 (recipe beam), 1 GPU"}}), flush=True)
