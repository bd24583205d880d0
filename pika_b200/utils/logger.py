"""Running-loss log lines for the training driver.

Public contract (what the trainer script calls, and what log-scraping tools of the reference recipe parse):
``Logger(file, every_n_labels, tags[, divisors])``, ``update_and_log(n_labels, [loss per tag])``,
``summarize_and_log() -> (sum of the first loss, total labels)``; a window line reads
``<tag>: <avg> \\t...fps: <k labels/s> k`` and the epoch line ``Finished, Overall Avg <tag>: <avg>\\t...Avg fps:<..> k``.
The bookkeeping is two accumulators (window, epoch) instead of parallel lists.
"""
import time


class _Acc:
    """label count + per-tag loss sums since ``t0``"""

    def __init__(self, k):
        self.clear(k)

    def clear(self, k):
        self.labels, self.sums, self.t0 = 0, [0.0] * k, time.time()

    def add(self, n, losses):
        self.labels += n
        self.sums = [s + float(l) for s, l in zip(self.sums, losses)]

    def averages(self, div):
        return [s / d / float(self.labels) for s, d in zip(self.sums, div)]

    def rate_k(self):
        return self.labels / (time.time() - self.t0) / 1000


class Logger(object):
    def __init__(self, log_file, log_per_nframes, tags, loss_per_frame=(1.0,)):
        self.log_file, self.log_per_nframes, self.tags = log_file, log_per_nframes, list(tags)
        k = len(self.tags)
        self.loss_per_frame = list(loss_per_frame) if len(loss_per_frame) == k else [1.0] * k
        self._window, self._epoch = _Acc(k), _Acc(k)

    def update_and_log(self, num_frames, loss):
        self._window.add(num_frames, loss)
        self._epoch.add(num_frames, loss)
        if self._window.labels < self.log_per_nframes:
            return
        cells = ''.join('{}: {:.3f} \t'.format(t, a) for t, a in zip(self.tags, self._window.averages(self.loss_per_frame)))
        self.log_file.write(cells + 'fps: {:.6f} k\n'.format(self._window.rate_k()))
        self.log_file.flush()
        self._window.clear(len(self.tags))

    def summarize_and_log(self):
        cells = ''.join('Finished, Overall Avg {}: {:.3f}\t'.format(t, a)
                        for t, a in zip(self.tags, self._epoch.averages(self.loss_per_frame)))
        self.log_file.write(cells + 'Avg fps:{:.6f} k\n'.format(self._epoch.rate_k()))
        return self._epoch.sums[0], self._epoch.labels
