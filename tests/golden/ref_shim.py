"""Import shims that let the UNMODIFIED reference Python run in the build container.

Only used by tests/golden/make_golden.py (which runs where /root/reference exists).
Nothing here is read on the GPU box.

Shims (all are environment repairs, none changes reference arithmetic):
  * sys.path: /root/reference (packages ``trainer``, ``decoder``, ``loader``, ``utils``).
  * kaldi / soundfile / resampy / warp_rnnt / editdistance: absent third-party packages are
    replaced by MagicMock modules so that ``import`` statements succeed; none of their
    functions is called by the code we execute.
  * numpy 2 removed ``np.sctypes`` (loader/audio.py:569,589): restored with the numpy-1 value.
  * torch >= 1.5 made ``LongTensor / int`` true division; decoder/beam_transducer.py:125
    (``prev_k = best_scores_id / num_words``) relies on the old floor semantics.  The module is
    loaded from source with that one expression rewritten to ``//``.
"""
import importlib.util
import sys
import types
from unittest import mock

import numpy as np

REF = "/root/reference"


def install():
    if REF not in sys.path:
        sys.path.insert(0, REF)
    for name in ["kaldi", "kaldi.util", "kaldi.util.table", "kaldi.util.options", "kaldi.util.io",
                 "kaldi.matrix", "kaldi.feat", "kaldi.feat.fbank", "kaldi.feat.mfcc", "kaldi.fstext",
                 "soundfile", "resampy", "warp_rnnt", "editdistance"]:
        if name not in sys.modules:
            sys.modules[name] = mock.MagicMock()
    if not hasattr(np, "sctypes"):
        np.sctypes = {"int": [np.int8, np.int16, np.int32, np.int64],
                      "uint": [np.uint8, np.uint16, np.uint32, np.uint64],
                      "float": [np.float16, np.float32, np.float64],
                      "complex": [np.complex64, np.complex128],
                      "others": [bool, object, bytes, str, np.void]}


def load_beam_module():
    """decoder.beam_transducer with the legacy floor-division restored (see module docstring)."""
    install()
    path = REF + "/decoder/beam_transducer.py"
    src = open(path).read()
    old = "prev_k = best_scores_id / num_words"
    assert src.count(old) == 1
    src = src.replace(old, "prev_k = best_scores_id // num_words")
    mod = types.ModuleType("decoder.beam_transducer")
    mod.__file__ = path
    exec(compile(src, path, "exec"), mod.__dict__)
    import decoder  # noqa: F401  (namespace package from /root/reference)
    sys.modules["decoder.beam_transducer"] = mod
    return mod


def model_args(output_dim, embd_dim=100, rnn_size=1024, dec_layers=2, dropout=0.2):
    """The argparse namespace trainer/model/transducer.py:27-72 reads (egs values)."""
    return types.SimpleNamespace(rnn_size=rnn_size, local_rank=0, decoder_type="rnn", brnn=True,
                                 encoder_type="transformer", embd_dim=embd_dim,
                                 padding_idx=output_dim, dropout=dropout, dec_layers=dec_layers,
                                 enc_layers=9)
