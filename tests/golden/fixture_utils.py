"""Deterministic weight tweaks shared by make_golden.py (applied to the REFERENCE model) and the
parity tests (applied to the drop-in model).  Works on any module exposing the reference's
attribute names (encoder.fc_out, embed, decoder, fc1, fc_gate, fc2)."""
import torch


def decode_fixture_reinit(m, blank_bias=5.0, s_l=0.15, s_hh=0.04, s_j=0.05, s_o=0.1, enc_scale=8.0, seed=2024, pred_scale=1.0):
    """A randomly initialised transducer never emits blank and its prediction net barely reacts to
    its input, which makes beam search degenerate.  Re-draw the prediction
    net / joint weights from wider seeded normals and bias blank so that hypotheses mix blanks,
    repeated and distinct tokens, pruning and both termination rules."""
    gg = torch.Generator().manual_seed(seed)
    with torch.no_grad():
        m.encoder.fc_out.weight *= enc_scale
        m.embed.weight.normal_(0, 1, generator=gg)
        for n, p in m.decoder.named_parameters():
            if "weight_hh" in n:
                # contractive recurrence (gain ~ s_hh * sqrt(H) / 4 < 1): a chaotic LSTM would amplify 1e-7
                # rounding differences between any two fp32 implementations into different hypotheses
                p.normal_(0, s_hh, generator=gg)
            elif "weight" in n:
                p.normal_(0, s_l, generator=gg)
            else:
                p.zero_()
        for l in (m.fc1, m.fc_gate):
            l.weight.normal_(0, s_j, generator=gg)
        m.fc2.weight.normal_(0, s_o, generator=gg)
        m.fc2.bias[0] += blank_bias
        if pred_scale != 1.0:
            # weaken the prediction net's pull on the joint (its columns of fc1 / fc_gate): at V = 6000 the default draw is
            # bistable -- all blanks, or a label run-away that ends at max_len with score gaps of 1e-3
            H = m.fc1.weight.shape[1] // 2
            m.fc1.weight[:, H:] *= pred_scale
            m.fc_gate.weight[:, H:] *= pred_scale
    return m


def decode_fixture_reinit_xf(m, blank_bias=2.5, s_j=0.05, s_o=0.1, out_scale=6.0, seed=2025):
    """Same purpose for the transformer prediction net (decoder_type='transformer'): wider embeddings and output projection so
    the prediction net reacts to its history, wider joint weights, a blank bias.  Conv / attention / FFN weights keep their
    seeded default initialisation."""
    gg = torch.Generator().manual_seed(seed)
    with torch.no_grad():
        m.embed.weight.normal_(0, 1, generator=gg)
        m.decoder.linear_out.weight *= out_scale
        for l in (m.fc1, m.fc_gate):
            l.weight.normal_(0, s_j, generator=gg)
        m.fc2.weight.normal_(0, s_o, generator=gg)
        m.fc2.bias[0] += blank_bias
    return m


def grad_fingerprint(g, n=384):
    """Compact, position-sensitive summary of a gradient tensor: [sum, abs-sum, l2 norm] followed by n strided samples
    (all elements when the tensor has at most n).  Used to pin whole-model gradients without storing 90 M floats."""
    import numpy as np
    f = g.detach().double().flatten()
    step = max(1, f.numel() // n)
    samp = f[::step][:n]
    head = np.array([float(f.sum()), float(f.abs().sum()), float(f.norm())])
    return np.concatenate([head, samp.numpy()])
