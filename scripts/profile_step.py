"""Kernel-time table of the bench train step via torch.profiler (CUPTI), near-native speed."""
import sys, os, types, collections, re
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import bench
from pika_b200 import engine
from pika_b200.frontend import FbankOptions, Frontend
from pika_b200.model.transducer import Net
from pika_b200.trainer.bmuf import BmufTrainer
from pika_b200.trainer.flat import FlatParams, SgdNesterovClip
from pika_b200.trainer.step import TrainStep
from pika_b200.utils.spec_augment import SpecAugment
a = types.SimpleNamespace(batch=int(os.environ.get("B", 32)), T=1000, U=150, V=6000)
dev = torch.device("cuda", 0)
ta = bench.train_args()
torch.manual_seed(777)
model = Net(bench.model_args(a.V), 240, a.V).to(dev); model.train()
flat = FlatParams(model); bmuf = BmufTrainer(0, 0, 1, model, 0.9, 1.0, flat=flat)
opt = SgdNesterovClip(flat, ta.initial_lr, 0.9, 3.0)
fe = Frontend(FbankOptions(num_mel_bins=80, low_freq=40.0, high_freq=-200.0, dither=0.0, window_type="hamming"), 1, 1, dev)
step = TrainStep(model, ta, fe, bmuf, opt, spec_augmentor=SpecAugment(15, 35))
B = a.batch
pcm = torch.from_numpy(bench.synth_pcm(B, a.T, 777)).to(dev)
n = pcm.shape[1]
new_len, frames = Frontend.lengths([n] * B, [1.0] * B)
i32 = lambda v: torch.tensor(v, dtype=torch.int32, device=dev)
batch = dict(pcm=pcm, target=torch.randint(1, a.V, (B, a.U), device=dev), n_samples=i32([n] * B), new_len=i32(new_len),
             n_frames=i32(frames), ali_lens=i32([a.U] * B), rate=torch.ones(B, device=dev), target_db=torch.full((B,), -25.0, device=dev),
             t_max=max(frames))
for _ in range(3):
    step(batch)
torch.cuda.synchronize()
from torch.profiler import profile, ProfilerActivity
with profile(activities=[ProfilerActivity.CUDA, ProfilerActivity.CPU]) as prof:
    step(batch); step(batch)
    torch.cuda.synchronize()
agg = collections.defaultdict(lambda: [0, 0.0])
tot = 0.0
for ev in prof.events():
    if ev.device_type == torch.autograd.DeviceType.CUDA:
        name = re.sub(r"\(.*", "", ev.name); name = name.replace("void ", "").replace("pk::", "")
        agg[name][0] += 1; agg[name][1] += ev.device_time_total if hasattr(ev, "device_time_total") else ev.cuda_time_total
        tot += agg[name][1] * 0
tot = sum(v[1] for v in agg.values())
print("total kernel us per step: %.1f" % (tot / 2))
for k, (c, t) in sorted(agg.items(), key=lambda x: -x[1][1])[:45]:
    print("%9.1f us %5.1f%% n=%5d  %s" % (t / 2, 100 * t / tot, c // 2, k[:100]))

# ---- idle gaps on the stream: time between the end of one kernel and the start of the next (two profiled steps)
evs = []
for ev in prof.events():
    if ev.device_type == torch.autograd.DeviceType.CUDA:
        tr = ev.time_range
        evs.append((tr.start, tr.end, re.sub(r"\(.*", "", ev.name).replace("void ", "").replace("pk::", "")[:60]))
evs.sort()
gaps = collections.defaultdict(lambda: [0, 0.0])
tot_gap, span = 0.0, (evs[-1][1] - evs[0][0]) if evs else 0.0
for (s0, e0, n0), (s1, e1, n1) in zip(evs, evs[1:]):
    g_ = s1 - e0
    if g_ > 1.0:
        tot_gap += g_
        gaps[(n0, n1)][0] += 1
        gaps[(n0, n1)][1] += g_
print("\nstream span %.1f us for 2 steps, idle between kernels %.1f us (%.1f us per step)" % (span, tot_gap, tot_gap / 2))
for (n0, n1), (c, t) in sorted(gaps.items(), key=lambda x: -x[1][1])[:25]:
    print("%8.1f us n=%4d  after %-45s before %s" % (t / 2, c // 2 or c, n0[:45], n1[:45]))
