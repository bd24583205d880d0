"""One RNN-T training batch on the GPU -- the body of run_one_epoch's hot loop
(trainer/train_transducer_bmuf_otfaug.py:71-130) with every numerically heavy stage replaced by the
sm_100a kernels: H2D of raw PCM -> on-GPU augmentation + fbank + splice + CMN/CMVN + SpecAugment ->
encoder / prediction net / fused joint+loss forward and backward -> inf-norm clip + Nesterov SGD ->
BMUF block sync every ``sync_period`` batches.
"""

from .. import engine
from .flat import lr_at
from .bmuf import SUCCESS


def encoder_out_lens(lens, lctx, rctx, stride):
    """T' = ceil((T - lctx - rctx) / stride) (trainer/train_transducer_bmuf_otfaug.py:79-82)"""
    l = lens - lctx - rctx
    return l // stride + (l % stride != 0).to(l.dtype)


class TrainStep:
    def __init__(self, model, args, frontend, bmuf, optimizer, offset=None, scale=None, spec_augmentor=None):
        self.model, self.args, self.frontend, self.bmuf, self.opt = model, args, frontend, bmuf, optimizer
        self.offset, self.scale, self.spec = offset, scale, spec_augmentor
        self.num_done = 0

    def features(self, batch):
        """batch: dict of device tensors (pcm int16 [B,n], n_samples, rate, target_db, new_len, n_frames) + t_max"""
        a = self.args
        sa = (0, 0, 0, 0)
        if self.spec is not None:
            sa = self.spec.draw(batch["t_max"], self.frontend.D)
        # the reference subtracts the utterance mean only inside ``if args.cmvn_stats:`` (:86-91), so --cmn without
        # --cmvn_stats is a no-op there; ``cmn_without_stats`` lets a caller that has no stats file (bench.py) keep CMN on
        cmn = bool(a.cmn) and (self.offset is not None or bool(getattr(a, "cmn_without_stats", False)))
        return self.frontend(batch["pcm"], batch["n_samples"], batch["rate"], batch["target_db"], batch["new_len"],
                             batch["n_frames"], batch["t_max"], out_dtype=engine.act_dtype(), cmn=cmn,
                             offset=self.offset, scale=self.scale, specaug=sa)

    def __call__(self, batch):
        """-> per-utterance costs [B] (device).  Mirrors :71-123 of the reference trainer."""
        a = self.args
        self.opt.flat.zero_grad()                                         # optimizer.zero_grad()
        feats = self.features(batch)
        len_batch = encoder_out_lens(batch["n_frames"], a.model_lctx, a.model_rctx, a.model_stride)
        costs = engine.transducer_loss(self.model, feats, batch["target"], len_batch, batch["ali_lens"])
        engine.assume_unit_loss_grad(True)                                # loss = costs.sum() (:99): upstream gradient is exactly 1
        try:
            costs.sum().backward()
        finally:
            engine.assume_unit_loss_grad(False)
        self.opt.step()                                                   # clip_grad_norm_(inf) + SGD(nesterov)
        self.end_of_item()
        return costs

    def skip(self):
        """An empty loader item (every utterance filtered, :100-101): no forward / backward / optimiser step, but the item
        still counts and still takes part in the periodic block sync (:112-123) -- a rank that skipped the collective
        while its peers entered it would pair their all-reduce with a later one."""
        self.end_of_item()

    def end_of_item(self):
        """``if num_done != 0 and num_done % sync_period == 0`` of the reference loop (:112-123): BMUF sync, new learning rate,
        fresh momentum buffer.  Runs for EVERY loader item, with or without data."""
        a = self.args
        if self.num_done != 0 and self.num_done % a.sync_period == 0:
            if self.bmuf.update_and_sync() != SUCCESS:
                raise FloatingPointError("BMUF: non-finite block delta")
            self.opt.reset(lr_at(a.initial_lr, a.final_lr, a.epoch * a.num_batches_per_epoch + self.num_done,
                                 a.num_epochs * a.num_batches_per_epoch))
        self.num_done += 1
