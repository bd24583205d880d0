"""Kernel-time table of one decode_batch of the bench decode workload (config 5: beam 16 x batch 64 x T=1500) via torch.profiler (CUPTI)."""
import sys, os, types, collections, re
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import bench
from pika_b200 import engine
from pika_b200.decoder.beam_transducer import GlobalScorer
from pika_b200.decoder.transducer_decoder import TransducerDecoder
a = types.SimpleNamespace(batch=int(os.environ.get("B", 64)), T=1500, U=150, V=6000, beam=16)
dev = torch.device("cuda", 0)
engine.set_precision("bf16")
model = bench.decode_model(a, dev)
dargs = types.SimpleNamespace(las_rescorer=None, las_rescorer_bw=None, bilas_rescorer=None, nonblk_reward=0.0)
dec = TransducerDecoder(model, a.batch, a.beam, n_best=1, blk=0, global_scorer=GlobalScorer(), cuda=True, beam_prune=True, args=dargs)
x = torch.from_numpy(bench.decode_feats(a.batch, a.T, 1)).to(dev)
tl = torch.full((a.batch,), bench.tprime(a.T), dtype=torch.int32)
ml = [int(t) + 100 for t in tl]
for _ in range(2):
    dec.decode_batch(x, tl, ml)
torch.cuda.synchronize()
from torch.profiler import profile, ProfilerActivity
with profile(activities=[ProfilerActivity.CUDA, ProfilerActivity.CPU]) as prof:
    ret, _ = dec.decode_batch(x, tl, ml)
    torch.cuda.synchronize()
steps = max(len(h[0]) for h in ret["predictions"]) + 1
agg = collections.defaultdict(lambda: [0, 0.0])
evs = []
for ev in prof.events():
    if ev.device_type == torch.autograd.DeviceType.CUDA:
        name = re.sub(r"\(.*", "", ev.name).replace("void ", "").replace("pk::", "")
        agg[name][0] += 1; agg[name][1] += ev.device_time_total
        evs.append((ev.time_range.start, ev.time_range.end))
tot = sum(v[1] for v in agg.values())
evs.sort()
span = evs[-1][1] - evs[0][0]
print("decode_batch: %d beam steps (longest hypothesis), %d graph replays, kernels per replay %d" % (steps, dec.last_replays, dec.kernels_per_replay))
print("total kernel us %.1f, stream span us %.1f (idle %.1f)" % (tot, span, span - tot))
for k, (c, t) in sorted(agg.items(), key=lambda x: -x[1][1])[:30]:
    print("%9.1f us %5.1f%% n=%5d  avg %7.2f us  %s" % (t, 100 * t / tot, c, t / c, k[:90]))
