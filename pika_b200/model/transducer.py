"""Generic transducer -- drop-in for trainer/model/transducer.py (reference).

``Net(opt, input_dim, output_dim)`` / ``forward(x, y, x_len, softmax)`` and the sub-module names
``encoder, embed, decoder, fc1, fc_gate, fc2`` (reached into by the decoder and the MBR trainer)
are the reference's.  TDNN-Transformer encoder (``encoder_type != 'rnn'``) with either prediction net of the reference:
the LSTM stack (``decoder_type == 'rnn'``) or the convolutional transformer (``'transformer'``,
trainer/model/rnnt_conv_transformer_lm.py).
"""
import torch.nn as nn

from .rnnt_conv_transformer_lm import Net as decoder_transformer
from .rnnt_tdnn_transformer import Net as encoder_tdnn


class Net(nn.Module):
    def __init__(self, opt, input_dim, output_dim):
        super().__init__()
        self.input_dim = input_dim
        self.output_dim = output_dim
        self.hid_dim = opt.rnn_size
        self.local_rank = getattr(opt, "local_rank", 0)
        self.decoder_type = opt.decoder_type
        if opt.encoder_type == "rnn":
            raise NotImplementedError("pika_b200: only the TDNN-Transformer encoder is on the hot path")
        self.encoder = encoder_tdnn(input_dim=input_dim, input_ctx=0, output_dim=self.hid_dim,
                                    tdnn_nhid=1024, tdnn_layers=9)
        self.pack_seq = False
        self.embed = nn.Embedding(output_dim + 1, opt.embd_dim, padding_idx=opt.padding_idx)
        if opt.decoder_type == "rnn":
            self.decoder = nn.LSTM(input_size=opt.embd_dim, hidden_size=self.hid_dim, dropout=opt.dropout,
                                   num_layers=opt.dec_layers, bidirectional=False, batch_first=True)
        else:                                          # trainer/model/transducer.py:62-68
            self.decoder = decoder_transformer(embeddings=self.embed, output_dim=self.hid_dim, d_model=512,
                                               num_layers=opt.dec_layers, heads=8, d_ff=2048, dropout=opt.dropout)
        self.fc1 = nn.Linear(2 * self.hid_dim, self.hid_dim)
        self.fc_gate = nn.Linear(2 * self.hid_dim, self.hid_dim)
        self.fc2 = nn.Linear(self.hid_dim, output_dim)

    def forward(self, x, y, x_len=None, softmax=True):
        """x [B,T,D] f32, y [B,U] int64 -> [B,T',U+1,V] log-probs (or logits if softmax=False)."""
        from pika_b200 import engine
        return engine.transducer_forward(self, x, y, softmax)

    def clean_hidden(self):
        """interface kept from the reference"""

    def reset_hidden(self, h, reset_idx):
        """interface kept from the reference"""
