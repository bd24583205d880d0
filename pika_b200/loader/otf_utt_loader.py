"""On-the-fly utterance loader -- drop-in for loader/otf_utt_loader.py (reference).

Module-level plugin API kept: ``register(parser)``, ``get_inputdim(args)``, ``dataloader(data_lst, rir, noise, args)``.
What moved: the per-utterance CPU work of the reference's producer threads (AudioSegment speed/gain
augmentation, PyKaldi fbank, splice; loader/otf_utt_loader.py:218-250) now runs on the GPU inside
``pika_b200.frontend.Frontend``.  The producer threads here only read raw int16 PCM from the ``.seq`` shards,
draw the augmentation parameters from the SAME random streams in the SAME order as the reference
(``random.randint`` for the speed rate, ``numpy.random.uniform`` for the gain target, :221-223), apply the TU
filter (:247) from the frame count the front end will produce, and assemble padded batches.

``dataloader`` yields, like the reference (:272-289), 4-tuples ``(data, target, lens, ali_lens)``; ``data`` is
either the reference's float feature tensor [B,Tmax,D] (``args.raw_batches`` false: features computed on the GPU
and returned as a CPU tensor) or, for the fused trainer path, a dict of raw-PCM tensors for ``TrainStep``.
End of stream = one ``None`` per worker, as in the reference.
"""
import queue
from random import randint
from threading import Thread

import numpy as np
import torch

from ..frontend import FbankOptions, Frontend
from . import kaldi_io


def get_inputdim(args):
    """full input dimension after splicing"""
    return args.feats_dim * (args.lctx + 1 + args.rctx)


def register(parser):
    """loader flags (same names, defaults and help strings as the reference)"""
    parser.add_argument('--lctx', type=int, default=10, help='left context for splice')
    parser.add_argument('--rctx', type=int, default=10, help='right context for splice')
    parser.add_argument('--max_len', type=int, default=6000, help='max length allowed to be loaded')
    parser.add_argument('--num_workers', type=int, default=5, help='number of workers to load/process')
    parser.add_argument('--sample_rate', type=int, default=16000, help='sample rate of waves')
    parser.add_argument('--buffer_size', type=int, default=128 * 1024, help='buffer size used to shuffle data')
    parser.add_argument('--batch_first', action='store_true', help='1st dim is batch or frame')
    parser.add_argument('--reverse_labels', action='store_true', help='reverse labels for training, eg for LAS')
    parser.add_argument('--feat_config', type=str, default=None, help='feature extraction config file')
    parser.add_argument('--stride', type=int, default=1, help='strides for subsampling input')
    parser.add_argument('--batch_size', type=int, default=1024, help='batch size')
    parser.add_argument('--SOS', type=int, default=-1, help='start of seq id, valid when beyond 0')
    parser.add_argument('--EOS', type=int, default=-1, help='end of seq id, valid when beyond 0')
    parser.add_argument('--queue_size', type=int, default=8, help='queue size for threading')
    parser.add_argument('--TU_limit', type=int, default=15000,
                        help='limits on the product of T (utt length) and U (label length) to avoid GPU OOM')
    parser.add_argument('--padding_tgt', type=int, default=-1, help='padding index for targets')
    parser.add_argument('--feats_dim', type=int, default=40, help='dimension of input feature (before splicing)')
    parser.add_argument('--snr_range', type=str, default='', help='comma separated SNR range in dB')
    parser.add_argument('--gain_range', type=str, default='55,10', help='comma separated negative gain range in dB')
    parser.add_argument('--speed_rate', type=str, default='0.9,1.0,1.1', help='comma separated rate for speed perturbation')
    parser.add_argument('--verbose', action='store_true', help='printing out warnings')


def put_thread(q, generator, *gen_args):
    for item in generator(*gen_args):
        q.put(item)
        if item is None:
            break


def otf_utt_generator(data_triplets, rir, noise, args):
    """raw-PCM batches for one worker; mirrors the control flow of loader/otf_utt_loader.py:165-299"""
    if args.stride != 1:
        raise NotImplementedError("pika_b200 loader: --stride 1 only (the recipes never subsample in the loader)")
    batch_size = args.batch_size
    speed_rate = [float(r) for r in args.speed_rate.split(',')]
    gain_lo, gain_hi = [-float(g) for g in args.gain_range.split(',')]
    pcm, tgt, meta = [], [], []
    batch_idx = 0
    for mrk_fn, seq_fn, ali_rspec in data_triplets:
        ali_reader = kaldi_io.read_int_vector_ark(ali_rspec)
        for (uttid, audio_np), (uttid1, ali) in zip(kaldi_io.iter_mrk_seq(mrk_fn, seq_fn), ali_reader):
            assert uttid == uttid1
            spr = speed_rate[randint(0, len(speed_rate) - 1)]
            target_db = np.random.uniform(gain_lo, gain_hi)
            ali = np.array(ali)
            if args.reverse_labels:
                ali = ali[::-1]
            if args.SOS >= 0:
                ali = np.concatenate(([args.SOS], ali))
            if args.EOS >= 0:
                ali = np.concatenate((ali, [args.EOS]))
            new_len, frames = Frontend.lengths([audio_np.shape[0]], [spr])
            utt_len = frames[0]
            if utt_len > 0 and utt_len <= args.max_len and ali.shape[0] * utt_len // 3 <= args.TU_limit:
                pcm.append(audio_np)
                tgt.append(ali.astype(np.int32))
                meta.append((audio_np.shape[0], spr, target_db, new_len[0], utt_len))
            batch_idx += 1
            if batch_idx == batch_size:
                yield assemble(pcm, tgt, meta, args)
                pcm, tgt, meta, batch_idx = [], [], [], 0
    yield None


def assemble(pcm, tgt, meta, args):
    """padded raw batch (or the reference's empty-batch tuple, loader/otf_utt_loader.py:283-287)"""
    if not pcm:
        return None, None, torch.IntTensor([0]), torch.IntTensor([0])
    B = len(pcm)
    n_max = max(max(m[0] for m in meta), max(m[3] for m in meta))
    u_max = max(len(t) for t in tgt)
    pin = torch.cuda.is_available()          # page-locked staging: the trainer's non_blocking H2D copy overlaps the previous step
    pcm_t = torch.zeros(B, n_max, dtype=torch.int16, pin_memory=pin)
    target = torch.full((B, u_max), args.padding_tgt, dtype=torch.int32, pin_memory=pin)
    for i in range(B):
        pcm_t[i, :meta[i][0]] = torch.from_numpy(pcm[i].copy())
        target[i, :len(tgt[i])] = torch.from_numpy(tgt[i])
    raw = dict(pcm=pcm_t,
               n_samples=torch.tensor([m[0] for m in meta], dtype=torch.int32),
               rate=torch.tensor([m[1] for m in meta], dtype=torch.float32),
               target_db=torch.tensor([m[2] for m in meta], dtype=torch.float32),
               new_len=torch.tensor([m[3] for m in meta], dtype=torch.int32),
               n_frames=torch.tensor([m[4] for m in meta], dtype=torch.int32),
               t_max=max(m[4] for m in meta))
    lens = raw["n_frames"].clone()
    ali_lens = torch.tensor([len(t) for t in tgt], dtype=torch.int32)
    return raw, target, lens, ali_lens


_frontends = {}


def _frontend_for(args, device):
    key = (args.feat_config, args.lctx, args.rctx, str(device))
    if key not in _frontends:
        opts = FbankOptions.from_config(args.feat_config) if args.feat_config else FbankOptions(num_mel_bins=args.feats_dim)
        if getattr(args, "no_dither", False):          # explicit opt-out (parity runs); otherwise the feature config decides
            opts.dither = 0.0
        _frontends[key] = Frontend(opts, args.lctx, args.rctx, device)
    return _frontends[key]


def raw_to_features(raw, args, device="cuda"):
    """reference-shaped float features [B,Tmax,D] on the CPU from a raw batch (GPU front end, then D2H)"""
    fe = _frontend_for(args, device)
    dev = {k: (v.to(device) if torch.is_tensor(v) else v) for k, v in raw.items()}
    out = fe(dev["pcm"], dev["n_samples"], dev["rate"], dev["target_db"], dev["new_len"], dev["n_frames"], dev["t_max"],
             out_dtype=torch.float32, cmn=False)
    return out.cpu()


def dataloader(data_lst, rir, noise, args):
    """
    Args:
        data_lst: list of mrk and seq of input audios, and label ark
        rir, noise: unused lists (as in the shipped reference recipes)
    """
    data_triplets = kaldi_io.read_lst(data_lst)
    num_per_worker = (len(data_triplets) + args.num_workers - 1) // args.num_workers
    lst = [data_triplets[i:i + num_per_worker] for i in range(0, len(data_triplets), num_per_worker)]
    assert len(lst) == args.num_workers
    q = queue.Queue(args.queue_size)
    threads = [Thread(target=put_thread, args=(q, otf_utt_generator, lst[i], rir, noise, args)) for i in range(args.num_workers)]
    for t in threads:
        t.daemon = True
        t.start()
    num_done = 0
    raw_mode = bool(getattr(args, "raw_batches", False))
    while True:
        item = q.get()
        if item is None:
            num_done += 1
            if num_done == args.num_workers:
                break
            continue
        raw, target, lens, ali_lens = item
        if raw is None or raw_mode:
            yield raw, target, lens, ali_lens
        else:
            data = raw_to_features(raw, args)
            if not args.batch_first:
                data, target = data.transpose(0, 1).contiguous(), target.t().contiguous()
            yield data, target, lens, ali_lens
    for t in threads:
        t.join()
