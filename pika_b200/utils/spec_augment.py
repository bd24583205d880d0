"""SpecAugment -- drop-in for utils/spec_augment.py (reference).

Same constructor / ``apply(inp)`` contract and, importantly, the same random streams in the same
order (torch ``Uniform.sample()`` for the spans, ``numpy.random.randint`` for the starts), so a run
seeded like the reference draws identical masks.  ``draw()`` exposes the draw so the fused GPU front
end can apply the masks inside its last kernel instead of two slice writes + two host syncs.
"""
import numpy as np
from torch.distributions.uniform import Uniform


class SpecAugment(object):
    def __init__(self, max_freq_span, max_time_span, batch_first=True):
        self.freq_span_sampler = Uniform(0.0, float(max_freq_span + 1))
        self.time_span_sampler = Uniform(0.0, float(max_time_span + 1))
        self.batch_first = batch_first

    def draw(self, num_frames, num_freq):
        """-> (freq_start, freq_span, time_start, time_span); spans of 0 mean 'no mask'
        (utils/spec_augment.py:13-20, including its ``randint(0, dim - span)`` ranges)."""
        freq_span = int(self.freq_span_sampler.sample().item())
        time_span = int(self.time_span_sampler.sample().item())
        freq_start = np.random.randint(0, num_freq - freq_span) if freq_span > 0 else 0
        time_start = np.random.randint(0, num_frames - time_span) if time_span > 0 else 0
        return int(freq_start), freq_span, int(time_start), time_span

    def apply(self, inp):
        """in place on a [batch, frame, freq] tensor; one mask pair shared by the whole batch"""
        f0, fs, t0, ts = self.draw(inp.size()[1], inp.size()[-1])
        if fs > 0:
            inp[:, :, f0:f0 + fs] = 0.0
        if ts > 0:
            inp[:, t0:t0 + ts, :] = 0.0
