"""Transducer training script -- drop-in for trainer/train_transducer_bmuf_otfaug.py (reference).

Same positional arguments and flags (argparse below mirrors :148-253 + the loader's ``register``), same
per-rank log / model file naming, same epoch structure (``run_one_epoch``: :32-145).  Launch exactly like the
reference recipe (one process per GPU, ``WORLD_SIZE`` / ``--local_rank`` from the launcher):

    python -m torch.distributed.run --nproc-per-node 8 --master-addr 127.0.0.1 -m pika_b200.trainer.train_transducer_bmuf_otfaug \\
        transducer data.WORKER-ID.lst log.WORKER-ID out/ --cuda --encoder_type transformer ... (flags of egs/train_transducer_bmuf_otfaug.sh)

What runs underneath: raw-PCM loader threads -> GPU front end -> sm_100a model / loss / optimiser kernels -> BMUF over NCCL.
"""
import argparse
import importlib
import os
import sys

import torch

from ..frontend import FbankOptions, Frontend
from ..loader import kaldi_io
from ..utils.logger import Logger
from ..utils.spec_augment import SpecAugment
from .. import engine
from .bmuf import BmufTrainer
from .flat import FlatParams, SgdNesterovClip, lr_at
from .step import TrainStep

MASTER_NODE = 0


def run_one_epoch(epoch, model, log_f, args, bmuf_trainer, training):
    """one epoch of training (trainer/train_transducer_bmuf_otfaug.py:32-145)"""
    log_f.write('===> Epoch {} <===\n'.format(epoch))
    total_num_batches = args.num_epochs * args.num_batches_per_epoch
    lr = lr_at(args.initial_lr, args.final_lr, epoch * args.num_batches_per_epoch, total_num_batches)
    log_f.write('===Using Learning Rate {}===\n'.format(lr))
    args.epoch = epoch
    optimizer = SgdNesterovClip(bmuf_trainer.flat, lr, args.momentum, args.grad_clip)
    loss_logger = Logger(args.log, args.log_per_n_frames, ['Loss'])
    spec = SpecAugment(args.max_freq_span, args.max_time_span) if args.spec_augment else None
    model.train(training)
    step = TrainStep(model, args, args.frontend, bmuf_trainer, optimizer, offset=args.offset, scale=args.scale, spec_augmentor=spec)
    dev = torch.device("cuda", args.local_rank)
    for num_done, (raw, target_cpu, len_cpu, ali_lens_cpu) in enumerate(args.dataloader(args.data_lst, args.rir, args.noise, args)):
        if raw is not None:
            batch = {k: (v.to(dev, non_blocking=True) if torch.is_tensor(v) else v) for k, v in raw.items()}
            batch["target"] = target_cpu.long().to(dev)
            batch["ali_lens"] = ali_lens_cpu.to(dev)
            if training:
                try:
                    costs = step(batch)
                except FloatingPointError:
                    return float('nan')                       # BMUF returned STOP (:113-114)
            else:
                with torch.no_grad():
                    feats = step.features(batch)
                    from .step import encoder_out_lens
                    tl = encoder_out_lens(batch["n_frames"], args.model_lctx, args.model_rctx, args.model_stride)
                    costs = engine.transducer_loss(model, feats, batch["target"], tl, batch["ali_lens"])
            loss = float(costs.sum().item())
        else:                                                 # empty batch (:100-101)
            loss = 0.0
            if training:
                try:
                    step.skip()                               # still counts, still syncs (:112-123)
                except FloatingPointError:
                    return float('nan')
        labels = int(ali_lens_cpu.sum().item())
        loss_logger.update_and_log(labels, [loss])
    if training and bmuf_trainer.update_and_sync() != 1:
        return float('nan')
    tot_loss, tot_num = loss_logger.summarize_and_log()
    loss_tensor = torch.tensor([tot_loss, float(tot_num)], dtype=torch.float32, device=dev)
    bmuf_trainer.sum_reduce(loss_tensor)                      # aggregate across workers (:140-145)
    bmuf_trainer.broadcast(loss_tensor)
    return (loss_tensor[0] / loss_tensor[1]).item()


def build_parser():
    parser = argparse.ArgumentParser(description='Transducer training')
    parser.add_argument('nnet_proto', type=str, help='pytorch NN proto definition filename')
    parser.add_argument('data_lst', type=str, help='list of mrk, seq, ali files for data')
    parser.add_argument('log', type=str, help='log file for the job')
    parser.add_argument('output_dir', type=str, help='path to save the final model')
    parser.add_argument('--init_model', type=str, default=None, help='initial model')
    parser.add_argument('--rir_lst', type=str, default=None, help='mrk and seq files for rir')
    parser.add_argument('--noise_lst', type=str, default=None, help='mrk and seq files for noise')
    parser.add_argument('--encoder_type', type=str, default='rnn', choices=['rnn', 'transformer'])
    parser.add_argument('--decoder_type', type=str, default='rnn', choices=['rnn', 'transformer'])
    parser.add_argument('--layers', type=int, default=-1)
    parser.add_argument('--enc_layers', type=int, default=2)
    parser.add_argument('--dec_layers', type=int, default=2)
    parser.add_argument('--rnn_size', type=int, default=512)
    parser.add_argument('--rnn_type', type=str, default='LSTM', choices=['LSTM'])
    parser.add_argument('--embd_dim', type=int, default=300)
    parser.add_argument('--output_dim', type=int, default=8000)
    parser.add_argument('--model_lctx', type=int, default=0)
    parser.add_argument('--model_rctx', type=int, default=0)
    parser.add_argument('--model_stride', type=int, default=1)
    parser.add_argument('--brnn', action="store_true")
    parser.add_argument('--cmn', action="store_true")
    parser.add_argument('--cmvn_stats', type=str, default=None)
    parser.add_argument('--optim', type=str, default='sgd', choices=['sgd', 'adam', 'adadelta'])
    parser.add_argument('--grad_clip', type=float, default=-1.0)
    parser.add_argument('--initial_lr', type=float, default=1.0)
    parser.add_argument('--final_lr', type=float, default=1.0)
    parser.add_argument('--momentum', type=float, default=0.9)
    parser.add_argument('--num_epochs', type=int, default=15)
    parser.add_argument('--num_batches_per_epoch', type=int, default=1000)
    parser.add_argument('--dropout', type=float, default=0.3)
    parser.add_argument('--padding_idx', type=int, default=-1)
    parser.add_argument('--loader', choices=['otf_utt'], default='otf_utt')
    parser.add_argument('--log_per_n_frames', type=int, default=1024 * 1024)
    parser.add_argument('--seed', type=int, default=777)
    parser.add_argument('--cuda', action='store_true')
    parser.add_argument('--local_rank', type=int, default=None)
    parser.add_argument('--block_momentum', type=float, default=0.9)
    parser.add_argument('--block_lr', type=float, default=1.0)
    parser.add_argument('--sync_period', type=int, default=100)
    parser.add_argument('--spec_augment', action='store_true')
    parser.add_argument('--max_freq_span', type=int, default=15)
    parser.add_argument('--max_time_span', type=int, default=35)
    parser.add_argument('--precision', choices=['bf16', 'fp32'], default='bf16', help='pika_b200: compute mode')
    return parser


def main(argv=None):
    parser = build_parser()
    args, _ = parser.parse_known_args(argv)
    loader_module = importlib.import_module('pika_b200.loader.' + args.loader + '_loader')
    loader_module.register(parser)
    args = parser.parse_args(argv)
    args.input_dim = loader_module.get_inputdim(args)
    args.dataloader = loader_module.dataloader
    args.raw_batches = True
    world_size = int(os.environ.get('WORLD_SIZE', '1'))
    if args.local_rank is None:
        args.local_rank = int(os.environ.get('LOCAL_RANK', '0'))
    assert args.cuda and torch.cuda.is_available(), "pika_b200 trains on the GPU (there is no CPU fallback)"
    torch.cuda.set_device(args.local_rank)
    dev = torch.device("cuda", args.local_rank)
    args.rir, args.noise = [], []
    args.data_lst = args.data_lst.replace('WORKER-ID', str(args.local_rank))
    args.log = args.log.replace('WORKER-ID', str(args.local_rank))
    log_f = open(args.log, 'w')
    args.log = log_f
    engine.set_precision(args.precision)
    engine.set_seed(args.seed + args.local_rank)
    nnet_module = importlib.import_module("pika_b200.model." + args.nnet_proto)
    torch.manual_seed(args.seed)
    if args.init_model is None:
        model = nnet_module.Net(args, args.input_dim, args.output_dim)
    else:
        model = torch.load(args.init_model, map_location=lambda storage, loc: storage, weights_only=False)
    model.to(dev)
    flat = FlatParams(model)
    bmuf_trainer = BmufTrainer(MASTER_NODE, args.local_rank, world_size, model, args.block_momentum, args.block_lr, flat=flat)
    num_param = sum(p.numel() for p in model.parameters())
    log_f.write('*' * 60 + '\n')
    log_f.write('model proto: {}\ninput  dim: {},\toutput dim: {},\nhidden dim: {},\tnum of enc_layers: {}\n'
                'num of dec_layers: {},\trnn_type: {}\nmodel size: {} M\n'.format(args.nnet_proto, args.input_dim, args.output_dim,
                                                                                args.rnn_size, args.enc_layers, args.dec_layers,
                                                                                args.rnn_type, num_param / 1000 / 1000))
    log_f.write('*' * 60 + '\n')
    log_f.flush()
    opts = FbankOptions.from_config(args.feat_config) if args.feat_config else FbankOptions(num_mel_bins=args.feats_dim)
    # opts.dither is honoured (egs/fbank.conf: dither=1): counter-based Gaussian dither in the fbank kernel; set dither=0 in the
    # feature config for bit-reproducible features (Kaldi's own RNG stream is not reproduced)
    args.frontend = Frontend(opts, args.lctx, args.rctx, dev)
    args.offset = args.scale = None
    if args.cmvn_stats:
        try:
            off, sc = kaldi_io.cmvn_offset_scale(args.cmvn_stats, args.lctx + args.rctx + 1)
        except ValueError as e:
            print(str(e))
            sys.exit()
        args.offset = torch.from_numpy(off).float().to(dev)
        args.scale = torch.from_numpy(sc).float().to(dev)
    for epoch in range(0, args.num_epochs):
        run_one_epoch(epoch, model, log_f, args, bmuf_trainer, True)
        current_model = '{}/model.epoch.{}.{}'.format(args.output_dir, epoch, args.local_rank)
        with open(current_model, 'wb') as f:
            torch.save(model, f)
    log_f.write('Training Finished')
    log_f.flush()


if __name__ == '__main__':
    main()
