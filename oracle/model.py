"""Oracle: TDNN-Transformer encoder, LSTM prediction net and gated joint, restated with
torch-CPU fp32 primitives over a plain ``state_dict`` (no reference classes).

TEST INFRASTRUCTURE ONLY (see oracle/__init__.py).  Pinned against the reference's own
``nn.Module``s by tests/golden/make_golden.py (run in the build container).

All paths relative to /root/reference.
"""
import math
import torch
import torch.nn.functional as F

TDNN_DIL_STRIDE = [(1, 1)] * 3 + [(3, 1)] * 5 + [(3, 4)]   # trainer/model/rnnt_tdnn_transformer.py:44-59
HEADS = [16, 16, 8]                                         # trainer/model/rnnt_tdnn_transformer.py:62
BN_EPS = 1e-5                                               # nn.BatchNorm1d default
LN_EPS = 1e-6                                               # trainer/model/modules/transformer.py:82


def batchnorm(x, sd, name, train):
    """nn.BatchNorm1d over rows of x [N, C].  train=True -> batch statistics (biased var),
    padded frames included (trainer/model/rnnt_tdnn_transformer.py:76-82)."""
    w, b = sd[name + ".weight"], sd[name + ".bias"]
    if train:
        mean = x.mean(0)
        var = x.var(0, unbiased=False)
    else:
        mean, var = sd[name + ".running_mean"], sd[name + ".running_var"]
    return (x - mean) * torch.rsqrt(var + BN_EPS) * w + b


def layernorm(x, sd, name):
    return F.layer_norm(x, (x.shape[-1],), sd[name + ".weight"], sd[name + ".bias"], LN_EPS)


def linear(x, sd, name):
    return F.linear(x, sd[name + ".weight"], sd[name + ".bias"])


def tdnn(x, sd, name, dil, stride):
    """nn.Conv2d(1, C, (3, C), dilation=(dil,1), stride=(stride,1)) on [B,1,T,C]
    == sum_k x[b, t*stride + k*dil, :] @ W[:, 0, k, :].T  (trainer/model/rnnt_tdnn_transformer.py:44-59, 81)."""
    w = sd[name + ".weight"]            # [C_out, 1, 3, C_in]
    b = sd[name + ".bias"]
    B, T, C = x.shape
    t_out = (T - 2 * dil - 1) // stride + 1
    idx = torch.arange(t_out) * stride
    y = b.view(1, 1, -1).expand(B, t_out, -1).clone()
    for k in range(3):
        y = y + x[:, idx + k * dil, :] @ w[:, 0, k, :].t()
    return y


def mha(x, sd, name, heads, mask=None):
    """MultiHeadedAttention.forward, self-attention, optional mask [B,T,T] (True = dropped key), no cache, no dropout
    (trainer/model/modules/multi_headed_attn.py:180-184, 199-207, 212-216, 220-223, 231-241)."""
    B, T, D = x.shape
    dh = D // heads

    def shape(z):
        return z.view(B, T, heads, dh).transpose(1, 2)

    k = shape(linear(x, sd, name + ".linear_keys"))
    v = shape(linear(x, sd, name + ".linear_values"))
    q = shape(linear(x, sd, name + ".linear_query"))
    q = q / math.sqrt(dh)                                  # scale BEFORE QK^T (:205)
    scores = torch.matmul(q, k.transpose(2, 3)).float()
    if mask is not None:
        scores = scores.masked_fill(mask.unsqueeze(1), -1e18)   # :214-216
    attn = torch.softmax(scores, dim=-1)
    ctx = torch.matmul(attn, v).transpose(1, 2).contiguous().view(B, T, D)
    return linear(ctx, sd, name + ".final_linear")


def transformer_layer(x, sd, name, heads, mask=None):
    """TransformerEncoderLayer.forward (trainer/model/modules/transformer.py:85-100) +
    PositionwiseFeedForward.forward (trainer/model/modules/position_ffn.py:27-39); dropout off."""
    h = mha(layernorm(x, sd, name + ".layer_norm"), sd, name + ".self_attn", heads, mask) + x
    ff = name + ".feed_forward"
    inter = F.relu(linear(layernorm(h, sd, ff + ".layer_norm"), sd, ff + ".w_1"))
    return linear(inter, sd, ff + ".w_2") + h


def encoder_forward(sd, x, train=True, prefix="encoder.", taps=None):
    """Net.forward (trainer/model/rnnt_tdnn_transformer.py:73-89).  x [B,T,D] -> [B,T',H].
    ``taps`` (dict) receives named intermediate activations when given."""
    B = x.shape[0]
    C = sd[prefix + "fc_in.weight"].shape[0]
    h = F.relu(linear(x, sd, prefix + "fc_in")).reshape(-1, C)
    h = batchnorm(h, sd, prefix + "bn_in", train).view(B, -1, C)
    if taps is not None:
        taps["bn_in"] = h
    for l, (dil, stride) in enumerate(TDNN_DIL_STRIDE):
        h = F.relu(tdnn(h, sd, prefix + "hidden_conv.%d" % l, dil, stride))
        T = h.shape[1]
        h = batchnorm(h.reshape(-1, C), sd, prefix + "hidden_bn.%d" % l, train).view(B, T, C)
        if taps is not None:
            taps["tdnn%d" % l] = h
        if (l + 1) % 3 == 0:
            h = transformer_layer(h, sd, prefix + "transformer.%d" % (l // 3), HEADS[l // 3])
            if taps is not None:
                taps["xf%d" % (l // 3)] = h
    T = h.shape[1]
    h = batchnorm(h.reshape(-1, C), sd, prefix + "bn_final", train)
    h = linear(h, sd, prefix + "fc_out")
    return h.view(B, T, -1)


def lstm_forward(sd, x, prefix="decoder.", layers=2, state=None):
    """nn.LSTM(batch_first=True), gate order i,f,g,o (torch convention), dropout off
    (trainer/model/transducer.py:56-61, 95).  x [B,U,E] -> (y [B,U,H], (h_n, c_n) [L,B,H])."""
    B, U, _ = x.shape
    hs, cs = [], []
    inp = x
    for l in range(layers):
        w_ih, w_hh = sd[prefix + "weight_ih_l%d" % l], sd[prefix + "weight_hh_l%d" % l]
        b = sd[prefix + "bias_ih_l%d" % l] + sd[prefix + "bias_hh_l%d" % l]
        H = w_hh.shape[1]
        if state is None:
            h = torch.zeros(B, H)
            c = torch.zeros(B, H)
        else:
            h, c = state[0][l], state[1][l]
        outs = []
        for u in range(U):
            g = inp[:, u] @ w_ih.t() + h @ w_hh.t() + b
            i, f, gg, o = g.chunk(4, dim=1)
            c = torch.sigmoid(f) * c + torch.sigmoid(i) * torch.tanh(gg)
            h = torch.sigmoid(o) * torch.tanh(c)
            outs.append(h)
        inp = torch.stack(outs, 1)
        hs.append(h)
        cs.append(c)
    return inp, (torch.stack(hs), torch.stack(cs))


def conv_transformer_lm_forward(sd, src, prefix="decoder.", heads=8):
    """Net.forward of the transformer prediction net (trainer/model/rnnt_conv_transformer_lm.py:59-80), dropout off.
    src [B,L] int64 -> [B,L,output_dim].  The embedding table is the transducer's ``embed`` (trainer/model/transducer.py:63);
    its padding_idx -1 resolves to the last row, whose id also marks padding keys (:66-68)."""
    emb_w = sd["embed.weight"]
    pad = emb_w.shape[0] - 1
    out = F.embedding(src, emb_w, padding_idx=pad)
    B, L = src.shape
    pad_mask = src.eq(pad).unsqueeze(1).expand(B, L, L)                                    # :66-68
    causal = torch.triu(torch.ones(L, L, dtype=torch.bool), diagonal=1).unsqueeze(0)      # :82-87
    mask = pad_mask | causal                                                               # :69-70
    l = 0
    while prefix + "conv.%d.weight" % l in sd:
        w, b = sd[prefix + "conv.%d.weight" % l], sd[prefix + "conv.%d.bias" % l]           # [N, C, 5]
        kw = w.shape[2]
        y = F.conv1d(out.transpose(1, 2), w, b, padding=kw - 1)[:, :, :-(kw - 1)]         # causal: drop the right overhang (:74)
        out = F.relu(y).transpose(1, 2)
        out = transformer_layer(out, sd, prefix + "transformer.%d" % l, heads, mask)       # :76
        l += 1
    return linear(layernorm(out, sd, prefix + "layer_norm"), sd, prefix + "linear_out")    # :77-78


def prednet_forward(sd, y, blank=0):
    """SOS prepend + embedding + LSTM (trainer/model/transducer.py:90-95), or the transformer prediction net when the
    state dict holds one (:96-97).  y [B,U] int64 -> [B,U+1,H]."""
    sos = torch.full((y.shape[0], 1), blank, dtype=torch.long)
    yy = torch.cat((sos, y.long()), 1)
    if "decoder.conv.0.weight" in sd:
        return conv_transformer_lm_forward(sd, yy)
    emb = F.embedding(yy, sd["embed.weight"], padding_idx=sd["embed.weight"].shape[0] - 1)   # padding row: no gradient
    return lstm_forward(sd, emb)[0]


def joint_forward(sd, enc, pred, softmax=True):
    """expand + cat + gated joint + log_softmax (trainer/model/transducer.py:96-111).
    enc [B,T,H], pred [B,U1,H] -> [B,T,U1,V]."""
    T, U1 = enc.shape[1], pred.shape[1]
    x = enc.unsqueeze(2).expand(-1, -1, U1, -1)
    y = pred.unsqueeze(1).expand(-1, T, -1, -1)
    z = torch.cat((x, y), -1)
    h = torch.tanh(linear(z, sd, "fc1")) * torch.sigmoid(linear(z, sd, "fc_gate"))
    out = linear(h, sd, "fc2")
    return F.log_softmax(out, -1) if softmax else out


def transducer_forward(sd, x, y, train=True, softmax=True):
    """Net.forward (trainer/model/transducer.py:74-112) without the hard-coded .cuda() at :91."""
    return joint_forward(sd, encoder_forward(sd, x, train), prednet_forward(sd, y), softmax)


def frame_lens_after_encoder(lens, lctx=21, rctx=21, stride=4):
    """trainer/train_transducer_bmuf_otfaug.py:79-82."""
    l = lens - lctx - rctx
    return l // stride + (l % stride != 0).to(l.dtype)
