"""``warp_rnnt``-compatible loss entry point (the reference imports ``from warp_rnnt import RNNTLoss``,
trainer/train_transducer_bmuf_otfaug.py:25, and calls ``RNNTLoss(blank=0, reduction='sum').apply``, :58)."""
from .engine import RNNTLossFn


class RNNTLoss(RNNTLossFn):
    """``RNNTLoss.apply(log_probs, labels, frame_lens, label_lens) -> costs [B]`` (blank = 0).
    Constructor kwargs are accepted and ignored exactly as in the reference's usage, where the instance's
    ``.apply`` is the autograd Function's classmethod."""

    def __init__(self, blank=0, reduction="sum"):
        super().__init__()
        assert blank == 0, "the reference path uses blank = 0"
