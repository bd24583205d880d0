"""Beam bookkeeping interface kept from decoder/beam_transducer.py (reference).  The per-utterance Python
class ``BeamMergeTransducer`` is replaced by the batched device kernel ``pk_beam_advance``
(pika_b200/csrc/beam.cu); only the scorer interface survives on the host."""


class GlobalScorer():
    """Global rescorer interface (identity, as in the reference: decoder/beam_transducer.py:246-258)."""

    def __init__(self):
        pass

    def score(self, beam, logprobs):
        return logprobs
