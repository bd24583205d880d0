"""CPU oracle for the RNN-T hot path of tencent-ailab/pika.

TEST INFRASTRUCTURE ONLY.  Nothing under ``pika_b200/`` may import this package;
only ``tests/``, ``__graft_entry__.smoke()`` and the ``cpu_baseline`` /
``--impl reference`` legs of ``bench.py`` use it, and there only as the checker
or as the timed CPU baseline -- never as the product path.

Every function restates one piece of the reference's algorithm on the CPU
(numpy for byte/integer/DP work, torch-CPU fp32 functional ops for the dense
layers, a small C file for the lattice DP at larger sizes) and cites the
reference ``file:line`` it follows (paths relative to ``/root/reference``).

Pinning status (see DESIGN.md "Oracle"):
  * model / joint / decoder / SpecAugment / BMUF update / MBR batch: pinned against
    the reference's own Python executed in the build container
    (``tests/golden/make_golden.py`` imports the modules from ``/root/reference`` --
    for the MBR trainer, which is a script, it executes the loop-body source text;
    for BMUF it runs the reference BmufTrainer on two gloo ranks -- and commits the
    outputs as ``tests/golden/*.npz``).
  * RNN-T loss: the reference delegates to the un-vendored, un-pinned
    ``warp_rnnt`` package (README.md:34-36).  Pinned against
    ``torchaudio.functional.rnnt_loss`` (an independent implementation of the
    same published recurrences) and a brute-force path enumeration.
  * fbank: the reference delegates to un-vendored PyKaldi/Kaldi
    (README.md:30-32).  Pinned against ``torchaudio.compliance.kaldi.fbank``
    (an independent port of Kaldi's feature-fbank) with ``dither=0``.
"""
