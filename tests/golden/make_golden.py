#!/usr/bin/env python
"""Generate the committed golden fixtures by EXECUTING THE REFERENCE's own Python modules
(imported from /root/reference through tests/golden/ref_shim.py) plus the two independent
third-party stand-ins for its un-vendored dependencies (torchaudio rnnt_loss / kaldi fbank).

Run in the build container only:   python tests/golden/make_golden.py
Outputs: tests/golden/*.npz (small; committed).  The GPU box never runs this script.
"""
import os
import sys

import numpy as np
import torch
import torch.nn.functional as F

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))
import ref_shim  # noqa: E402

ref_shim.install()
torch.set_num_threads(8)


def disable_dropout(model):
    """Parity fixtures run with dropout disabled (Philox streams cannot be reproduced by a custom
    kernel) while BatchNorm stays in train mode -- SURVEY.md section 7 'Dropout parity'."""
    for m in model.modules():
        if isinstance(m, torch.nn.Dropout):
            m.p = 0.0
    model.decoder.dropout = 0.0


def build_ref_model(V, seed=777, embd=100):
    from trainer.model.transducer import Net
    torch.manual_seed(seed)                       # trainer/train_transducer_bmuf_otfaug.py:295
    return Net(ref_shim.model_args(V, embd_dim=embd), 240, V)


def weight_fingerprint(model):
    out = {}
    for k, v in model.state_dict().items():
        if v.dtype.is_floating_point:
            out[k] = np.array([v.double().sum().item(), v.double().abs().sum().item(),
                               float(v.flatten()[0]), float(v.flatten()[-1])])
    return out


def ref_forward_cpu(m, x, y, softmax):
    """trainer/model/transducer.py:88-111 with the hard-coded SOS.cuda() (:91) left out."""
    enc = m.encoder(x)
    sos = torch.zeros(y.shape[0], 1).long()
    yy = torch.cat((sos, y), dim=1)
    pred, _ = m.decoder(m.embed(yy))
    T, U = enc.size(1), pred.size(1)
    xe = enc.unsqueeze(2).expand(-1, -1, U, -1)
    ye = pred.unsqueeze(1).expand(-1, T, -1, -1)
    out = torch.cat((xe, ye), dim=-1)
    out = m.fc2(torch.tanh(m.fc1(out)) * torch.sigmoid(m.fc_gate(out)))
    if softmax:
        out = F.log_softmax(out, dim=-1)
    return enc, pred, out


def golden_model():
    """Reference model forward + (torchaudio-loss) backward on a small seeded batch."""
    import torchaudio
    V, B, T, U = 40, 2, 120, 6
    m = build_ref_model(V)
    fp = weight_fingerprint(m)
    m.train()
    disable_dropout(m)
    g = torch.Generator().manual_seed(1234)
    x = torch.randn(B, T, 240, generator=g)
    y = torch.randint(1, V, (B, U), generator=g)
    lens = torch.tensor([T, T - 17], dtype=torch.int32)
    ulens = torch.tensor([U, U - 2], dtype=torch.int32)
    # intermediate activations via forward hooks
    taps = {}
    hooks = []
    for i in range(9):
        hooks.append(m.encoder.hidden_bn[i].register_forward_hook(
            lambda mod, inp, out, i=i: taps.__setitem__("tdnn%d" % i, out.detach().transpose(1, 2))))
    for i in range(3):
        hooks.append(m.encoder.transformer[i].register_forward_hook(
            lambda mod, inp, out, i=i: taps.__setitem__("xf%d" % i, out.detach())))
    enc, pred, logits = ref_forward_cpu(m, x, y, softmax=False)
    for h in hooks:
        h.remove()
    lp = F.log_softmax(logits, -1)
    tl = (lens - 42)
    tl = tl // 4 + (tl % 4 != 0).int()            # trainer/train_transducer_bmuf_otfaug.py:79-82
    costs = torchaudio.functional.rnnt_loss(lp, y.int(), tl, ulens, blank=0, reduction="none",
                                            fused_log_softmax=False)
    loss = costs.sum()
    loss.backward()
    out = dict(x=x.numpy(), y=y.numpy().astype(np.int32), lens=lens.numpy(), ulens=ulens.numpy(),
               tlens=tl.numpy(), enc=enc.detach().numpy(), pred=pred.detach().numpy(),
               logits=logits.detach().numpy(), costs=costs.detach().numpy(), V=np.array(V))
    for k, v in taps.items():
        out["tap_" + k] = v[:, ::7, ::13].contiguous().numpy()      # strided sample keeps the file small
    for k, v in fp.items():
        out["w_" + k] = v
    for k, p in m.named_parameters():
        gr = p.grad
        out["g_" + k] = np.concatenate([[gr.double().norm().item(), gr.double().abs().max().item()],
                                        gr.flatten()[:6].double().numpy()])
        from fixture_utils import grad_fingerprint
        out["gs_" + k] = grad_fingerprint(gr, 512)      # [sum, abs-sum, norm] + <= 512 strided samples: pins the DIRECTION, not just the norm
    # BatchNorm running stats after one train-mode forward
    for k, v in m.state_dict().items():
        if "running" in k:
            out["bn_" + k] = np.array([v.double().sum().item(), float(v.flatten()[0])])
    np.savez_compressed(os.path.join(HERE, "model_small.npz"), **out)
    print("model_small: costs", costs.tolist(), "enc", tuple(enc.shape), "logits", tuple(logits.shape))


from make_inputs import decode_big_inputs, full_shape_inputs  # noqa: E402  (shared with the GPU tests)


def golden_model_full():
    """The reference model at the BASELINE shape: T=1000 frames (T'=240), U=150, V=6000, B=2 with ragged lengths.
    Forward + torchaudio-loss backward on the CPU (about 10 GB, a few minutes); stores strided samples only."""
    import torchaudio
    from fixture_utils import grad_fingerprint
    V = 6000
    x_np, y_np, lens_np, ulens_np = full_shape_inputs(V=V)
    m = build_ref_model(V)
    m.train()
    disable_dropout(m)
    x, y = torch.from_numpy(x_np), torch.from_numpy(y_np)
    lens, ulens = torch.from_numpy(lens_np), torch.from_numpy(ulens_np)
    enc, pred, logits = ref_forward_cpu(m, x, y, softmax=False)
    logits.retain_grad()
    lp = F.log_softmax(logits, -1)
    tl = lens - 42
    tl = tl // 4 + (tl % 4 != 0).int()
    costs = torchaudio.functional.rnnt_loss(lp, y.int(), tl, ulens, blank=0, reduction="none", fused_log_softmax=False)
    costs.sum().backward()
    dl = logits.grad
    out = dict(seed=np.array(2025), tlens=tl.numpy(), costs=costs.detach().numpy(), V=np.array(V),
               enc=enc.detach()[:, ::5, ::7].contiguous().numpy(), pred=pred.detach()[:, ::3, ::7].contiguous().numpy(),
               logits=logits.detach()[:, ::9, ::7, ::53].contiguous().numpy(),
               lse=torch.logsumexp(logits.detach(), -1)[:, ::9, ::7].contiguous().numpy(),
               dlogits=dl[:, ::9, ::7, ::53].contiguous().numpy(),
               dlogits_blank=dl[:, ::3, ::3, 0].contiguous().numpy(),
               dlogits_colsum=dl.double().sum((0, 1, 2)).numpy())
    for k, p in m.named_parameters():
        out["gs_" + k] = grad_fingerprint(p.grad, 256)
    np.savez_compressed(os.path.join(HERE, "model_full_shape.npz"), **out)
    print("model_full_shape: costs", costs.tolist(), "logits", tuple(logits.shape))


from make_inputs import DECODE_BIG_REINIT  # noqa: E402


def golden_decode_big():
    """Reference decode_batch at the width of BASELINE config 5: beam 16, V=6000 (batch / frames reduced so the fixture
    generates in seconds on the CPU).  The encoder is replaced by a module that returns a seeded [B, T', H] tensor (both
    sides regenerate it from the seed): random fbank through a randomly initialised encoder gives nearly frame-independent
    outputs, on which beam search at V = 6000 either emits nothing or runs away with score gaps of 1e-3; frame-varying
    encoder outputs make every frame prefer different labels, with healthy margins.  Everything the beam loop executes
    (prediction net, joint, log-softmax, BeamMergeTransducer.advance, state reorder) is the reference's own code."""
    import types
    ref_shim.load_beam_module()
    from decoder.transducer_decoder import TransducerDecoder
    import decoder.beam_transducer as bt
    from fixture_utils import decode_fixture_reinit
    V, B, Tp, beam, nbest = 6000, 6, 72, 16, 4
    m = build_ref_model(V)
    m.eval()
    decode_fixture_reinit(m, **DECODE_BIG_REINIT)
    enc_np = decode_big_inputs(606, B, Tp)
    enc_t = torch.from_numpy(enc_np)

    class FixedEncoder(torch.nn.Module):
        def forward(self, x):
            return enc_t

    m.encoder = FixedEncoder()
    tl = torch.tensor([72, 72, 65, 60, 45, 34])
    dargs = types.SimpleNamespace(las_rescorer=None, las_rescorer_bw=None, bilas_rescorer=None, nonblk_reward=0.0)
    dec = TransducerDecoder(m, B, beam, n_best=nbest, blk=0, global_scorer=bt.GlobalScorer(), sm_scale=1.0, cuda=False,
                            beam_prune=True, args=dargs)
    with torch.no_grad():
        ret, enc = dec.decode_batch(torch.zeros(B, 1, 240), tl, max_len=[int(t) + 40 for t in tl])
    cases = {}
    for b in range(B):
        for n in range(nbest):
            cases["pred_%d_%d" % (b, n)] = np.array([int(t) for t in ret["predictions"][b][n]], np.int64)
            cases["score_%d_%d" % (b, n)] = np.array(float(ret["scores"][b][n]))
    print("decode_big", [len(cases["pred_%d_0" % b]) for b in range(B)], [float(ret["scores"][b][0]) for b in range(B)],
          "distinct tokens in best hyps:", len(set(int(t) for b in range(B) for t in cases["pred_%d_0" % b])))
    np.savez_compressed(os.path.join(HERE, "decode_big.npz"), seed=np.array(606), tlens=tl.numpy(),
                        dims=np.array([V, B, Tp, beam, nbest]), **cases)


def build_ref_model_xf(V, seed=778, embd=100, dec_layers=2):
    """reference transducer with the convolutional-transformer prediction net (trainer/model/transducer.py:62-68)"""
    from trainer.model.transducer import Net
    a = ref_shim.model_args(V, embd_dim=embd, dec_layers=dec_layers)
    a.decoder_type = "transformer"
    torch.manual_seed(seed)
    return Net(a, 240, V)


def xf_inputs(seed, B, Tp, H=1024):
    return np.random.default_rng(seed).standard_normal((B, Tp, H)).astype(np.float32)


def golden_model_xf():
    """Reference model with decoder_type='transformer': prediction net + joint forward and (torchaudio-loss) backward.  The
    encoder is replaced by seeded outputs (it is pinned by model_small.npz); everything from the embedding to the loss is the
    reference's own code.  The second utterance's label row ends in padding ids (= V, embed.padding_idx), which exercises the
    padding-key mask of trainer/model/rnnt_conv_transformer_lm.py:66-70."""
    import torchaudio
    V, B, Tp, U = 40, 3, 20, 9
    m = build_ref_model_xf(V)
    m.train()
    for mod in m.modules():
        if isinstance(mod, torch.nn.Dropout):
            mod.p = 0.0
    fp = weight_fingerprint(m)
    g = torch.Generator().manual_seed(99)
    y = torch.randint(1, V, (B, U), generator=g)
    ulens = torch.tensor([U, U - 3, U - 5], dtype=torch.int32)
    for b in range(B):
        y[b, int(ulens[b]):] = V
    enc = torch.from_numpy(xf_inputs(707, B, Tp)).requires_grad_(True)
    tl = torch.tensor([Tp, Tp - 4, Tp - 7], dtype=torch.int32)
    sos = torch.zeros(B, 1).long()
    pred = m.decoder(torch.cat((sos, y), dim=1))                       # trainer/model/transducer.py:96-97
    T_, U1 = enc.size(1), pred.size(1)
    out = torch.cat((enc.unsqueeze(2).expand(-1, -1, U1, -1), pred.unsqueeze(1).expand(-1, T_, -1, -1)), dim=-1)
    logits = m.fc2(torch.tanh(m.fc1(out)) * torch.sigmoid(m.fc_gate(out)))
    lp = F.log_softmax(logits, -1)
    yl = y.clone()
    yl[yl == V] = 0                                                     # label values past label_lens are never read by the loss
    costs = torchaudio.functional.rnnt_loss(lp, yl.int(), tl, ulens, blank=0, reduction="none", fused_log_softmax=False)
    costs.sum().backward()
    from fixture_utils import grad_fingerprint
    res = dict(y=y.numpy().astype(np.int64), ulens=ulens.numpy(), tlens=tl.numpy(), pred=pred.detach().numpy(),
               logits=logits.detach().numpy()[:, ::3, :, :], costs=costs.detach().numpy(), denc=grad_fingerprint(enc.grad, 512),
               dims=np.array([V, B, Tp, U]), seed=np.array(707))
    for k, v in fp.items():
        if not k.startswith("encoder."):
            res["w_" + k] = v
    for k, p in m.named_parameters():
        if not k.startswith("encoder."):
            res["gs_" + k] = grad_fingerprint(p.grad if p.grad is not None else torch.zeros_like(p), 512)
    np.savez_compressed(os.path.join(HERE, "model_xf.npz"), **res)
    print("model_xf: costs", costs.tolist(), "pred", tuple(pred.shape))


def golden_decode_xf():
    """Reference decode_batch with the transformer prediction net (decoder/transducer_decoder.py:117-120,151-171,195-200), beam 4
    and 8, on seeded encoder outputs.  The reference builds ``torch.cuda.LongTensor`` in that branch unconditionally (:166); for
    this CPU run the name is pointed at ``torch.LongTensor`` (an environment repair, no arithmetic involved)."""
    import types
    ref_shim.load_beam_module()
    from decoder.transducer_decoder import TransducerDecoder
    import decoder.beam_transducer as bt
    from fixture_utils import decode_fixture_reinit_xf
    V, B, Tp = 40, 4, 24
    m = build_ref_model_xf(V)
    m.eval()
    decode_fixture_reinit_xf(m)
    enc_t = torch.from_numpy(xf_inputs(808, B, Tp))

    class FixedEncoder(torch.nn.Module):
        def forward(self, x):
            return enc_t

    m.encoder = FixedEncoder()
    tl = torch.tensor([24, 21, 17, 9])
    cases = {}
    saved = torch.cuda.LongTensor
    torch.cuda.LongTensor = torch.LongTensor
    try:
        for name, beam, nbest in [("b4n2", 4, 2), ("b8n4", 8, 4)]:
            dargs = types.SimpleNamespace(las_rescorer=None, las_rescorer_bw=None, bilas_rescorer=None, nonblk_reward=0.0)
            dec = TransducerDecoder(m, B, beam, n_best=nbest, blk=0, global_scorer=bt.GlobalScorer(), sm_scale=1.0, cuda=False,
                                    beam_prune=True, args=dargs)
            with torch.no_grad():
                ret, _ = dec.decode_batch(torch.zeros(B, 1, 240), tl, max_len=[int(t) + 30 for t in tl])
            for b in range(B):
                for n in range(nbest):
                    cases["%s_pred_%d_%d" % (name, b, n)] = np.array([int(t) for t in ret["predictions"][b][n]], np.int64)
                    cases["%s_score_%d_%d" % (name, b, n)] = np.array(float(ret["scores"][b][n]))
            print("decode_xf", name, [len(cases["%s_pred_%d_0" % (name, b)]) for b in range(B)],
                  [cases["%s_pred_%d_0" % (name, b)].tolist() for b in range(2)], [float(ret["scores"][b][0]) for b in range(B)])
    finally:
        torch.cuda.LongTensor = saved
    np.savez_compressed(os.path.join(HERE, "decode_xf.npz"), seed=np.array(808), tlens=tl.numpy(), dims=np.array([V, B, Tp]), **cases)


class _FakeFst:
    """the slice of the kaldi.fstext VectorFst interface decoder/sorted_matcher.py uses, over an in-memory arc table"""

    class _W:
        def __init__(self, v):
            self.value = v

    class _Arc:
        def __init__(self, il, w, ns):
            self.ilabel, self.weight, self.nextstate = il, _FakeFst._W(w), ns

    class _Iter:
        def __init__(self, arcs):
            self.arcs, self.pos = arcs, 0

        def seek(self, i):
            self.pos = i

        def done(self):
            return self.pos >= len(self.arcs)

        def value(self):
            return _FakeFst._Arc(*self.arcs[self.pos])

    def __init__(self, arcs, finals):
        self._arcs, self._finals = arcs, finals

    def arcs(self, state):
        return _FakeFst._Iter(self._arcs[state])

    def final(self, state):
        return _FakeFst._W(self._finals[state])


def golden_decode_fst():
    """Reference decode_batch with on-the-fly FST shallow fusion (lm_scorer = the reference's own SortedMatcher,
    decoder/sorted_matcher.py, over a toy back-off LM held in an in-memory stand-in for the PyKaldi VectorFst): the V=40 model and
    inputs of decode_small, beam 4 and 8, lm_scorer_scale 0.5, nonblk_reward 0.45."""
    import types
    ref_shim.load_beam_module()
    from decoder.transducer_decoder import TransducerDecoder
    from decoder.sorted_matcher import SortedMatcher
    import decoder.beam_transducer as bt
    from fixture_utils import decode_fixture_reinit
    from make_inputs import toy_backoff_lm
    V = 40
    m = build_ref_model(V)
    m.eval()
    decode_fixture_reinit(m)
    d = np.load(os.path.join(HERE, "decode_small.npz"))
    x, tl = torch.from_numpy(d["x"]), torch.from_numpy(d["tlens"])
    arcs, finals = toy_backoff_lm(V)
    matcher = SortedMatcher(_FakeFst(arcs, finals), max(len(a) for a in arcs), V + 2, 1, [])
    cases = {}
    for name, beam, nbest in [("b4", 4, 2), ("b8", 8, 4)]:
        dargs = types.SimpleNamespace(las_rescorer=None, las_rescorer_bw=None, bilas_rescorer=None, nonblk_reward=0.45)
        dec = TransducerDecoder(m, 3, beam, n_best=nbest, blk=0, global_scorer=bt.GlobalScorer(), sm_scale=1.0, cuda=False,
                                lm_scorer=matcher, lm_scorer_scale=0.5, beam_prune=True, args=dargs)
        with torch.no_grad():
            ret, _ = dec.decode_batch(x, tl, max_len=[int(t) + 100 for t in tl])
        for b in range(3):
            for n in range(nbest):
                cases["%s_pred_%d_%d" % (name, b, n)] = np.array([int(t) for t in ret["predictions"][b][n]], np.int64)
                cases["%s_score_%d_%d" % (name, b, n)] = np.array(float(ret["scores"][b][n]))
        print("decode_fst", name, [[int(t) for t in cases["%s_pred_%d_0" % (name, b)] if t != 0] for b in range(3)],
              [round(float(cases["%s_score_%d_0" % (name, b)]), 3) for b in range(3)])
    np.savez_compressed(os.path.join(HERE, "decode_fst.npz"), **cases)


def golden_encoder_eval():
    """BASELINE config 1: encoder forward, 1 utterance, T=200, eval mode."""
    m = build_ref_model(40)
    m.eval()
    g = torch.Generator().manual_seed(99)
    x = torch.randn(1, 200, 240, generator=g)
    with torch.no_grad():
        enc = m.encoder(x)
    np.savez_compressed(os.path.join(HERE, "encoder_eval_T200.npz"), x=x.numpy(), enc=enc.numpy())
    print("encoder_eval_T200:", tuple(enc.shape))


def golden_decode():
    """Reference TransducerDecoder.decode_batch on CPU (beam search), eval mode."""
    import types
    ref_shim.load_beam_module()
    from decoder.transducer_decoder import TransducerDecoder
    import decoder.beam_transducer as bt
    V = 40
    m = build_ref_model(V)
    m.eval()
    from fixture_utils import decode_fixture_reinit
    decode_fixture_reinit(m)          # see fixture_utils.py: random init decodes degenerately
    g = torch.Generator().manual_seed(4321)
    B, T = 3, 130
    x = torch.randn(B, T, 240, generator=g)
    x_len_frames = torch.tensor([130, 118, 101])
    tl = x_len_frames - 42
    tl = tl // 4 + (tl % 4 != 0).long()
    cases = {}
    for name, beam, nbest, prune in [("b4n1", 4, 1, True), ("b4n4np", 4, 4, False), ("b8n2", 8, 2, True)]:
        dargs = types.SimpleNamespace(las_rescorer=None, las_rescorer_bw=None, bilas_rescorer=None,
                                      nonblk_reward=0.0)
        dec = TransducerDecoder(m, B, beam, n_best=nbest, blk=0, global_scorer=bt.GlobalScorer(),
                                sm_scale=1.0, cuda=False, beam_prune=prune, args=dargs)
        with torch.no_grad():
            ret, enc = dec.decode_batch(x, tl, max_len=[int(t) + 100 for t in tl])
        for b in range(B):
            for n in range(nbest):
                hyp = [int(t) for t in ret["predictions"][b][n]]
                cases["%s_pred_%d_%d" % (name, b, n)] = np.array(hyp, np.int64)
                cases["%s_score_%d_%d" % (name, b, n)] = np.array(float(ret["scores"][b][n]))
        print("decode", name, [len(cases["%s_pred_%d_0" % (name, b)]) for b in range(B)],
              [float(ret["scores"][b][0]) for b in range(B)])
    np.savez_compressed(os.path.join(HERE, "decode_small.npz"), x=x.numpy(), tlens=tl.numpy(),
                        enc=enc.numpy()[:, ::3, ::17], **cases)


def golden_specaug():
    """Reference SpecAugment (utils/spec_augment.py) with seeded torch + numpy RNGs."""
    from utils.spec_augment import SpecAugment
    out = {}
    for i, seed in enumerate([0, 1, 7, 777]):
        torch.manual_seed(seed)
        np.random.seed(seed)
        x = torch.ones(3, 200, 240)
        sa = SpecAugment(15, 35)
        sa.apply(x)
        sa.apply(x)                                 # two consecutive draws from the same streams
        out["mask_%d" % i] = np.packbits((x[0] == 0).numpy())
        out["seed_%d" % i] = np.array(seed)
    np.savez_compressed(os.path.join(HERE, "specaug.npz"), **out)
    print("specaug done")


def golden_frontend():
    """AudioSegment hot methods from the reference (loader/audio.py), splice from
    loader/otf_utt_loader.py, and torchaudio's Kaldi-compatible fbank as the PyKaldi stand-in."""
    import torchaudio
    from loader.audio import AudioSegment
    from loader.otf_utt_loader import splice
    rng = np.random.default_rng(5)
    out = {}
    n = 400 + 59 * 160 + 37
    # speech-like: sum of a few sinusoids + noise, int16
    t = np.arange(n) / 16000.0
    wav = 3000 * np.sin(2 * np.pi * 220 * t) + 1500 * np.sin(2 * np.pi * 1330 * t + 1.0) + \
        800 * rng.standard_normal(n)
    pcm = np.clip(np.round(wav), -32768, 32767).astype(np.int16)
    out["pcm"] = pcm
    for rate, db in [(0.9, -23.5), (1.0, -41.0), (1.1, -12.25)]:
        seg = AudioSegment(pcm, 16000)
        seg.change_speed(rate)
        seg.normalize(db)
        aug = seg._convert_samples_from_float32(seg._samples, "int16")
        key = "r%02d" % int(rate * 10)
        out["aug_" + key] = aug
        fb = torchaudio.compliance.kaldi.fbank(
            torch.from_numpy(aug.astype(np.float32)).unsqueeze(0), num_mel_bins=80,
            sample_frequency=16000.0, dither=0.0, low_freq=40.0, high_freq=-200.0,
            window_type="hamming", energy_floor=0.0)
        out["fbank_" + key] = fb.numpy()
        out["splice_" + key] = splice(fb.numpy(), 1, 1)[::5]
    np.savez_compressed(os.path.join(HERE, "frontend.npz"), **out)
    print("frontend done", {k: v.shape for k, v in out.items() if k.startswith("fbank")})


def golden_rnnt():
    """torchaudio rnnt_loss (independent implementation) on seeded cases incl. ragged lengths."""
    import torchaudio
    out = {}
    for i, (B, T, U, V) in enumerate([(1, 1, 0, 5), (2, 4, 3, 7), (3, 25, 9, 33), (4, 40, 17, 129)]):
        g = torch.Generator().manual_seed(100 + i)
        logits = 2.0 * torch.randn(B, T, U + 1, V, generator=g)
        labels = torch.randint(1, V, (B, max(U, 1)), generator=g).int()[:, :U]
        fl = torch.randint(max(1, T // 2), T + 1, (B,), generator=g).int()
        ll = torch.randint(U // 2, U + 1, (B,), generator=g).int()
        fl[0], ll[0] = T, U
        lp = F.log_softmax(logits, -1).requires_grad_(True)
        lab = labels if U > 0 else torch.zeros(B, 0, dtype=torch.int32)
        if U == 0:
            # torchaudio needs U>=1 storage; T=1,U=0 closed form: cost = -lp[0,0,blank]
            costs = -lp[:, 0, 0, 0]
        else:
            costs = torchaudio.functional.rnnt_loss(lp, lab, fl, ll, blank=0, reduction="none",
                                                    fused_log_softmax=False)
        costs.sum().backward()
        out["logits_%d" % i] = logits.numpy()
        out["labels_%d" % i] = lab.numpy()
        out["fl_%d" % i] = fl.numpy()
        out["ll_%d" % i] = ll.numpy()
        out["costs_%d" % i] = costs.detach().numpy()
        out["grads_%d" % i] = lp.grad.numpy()
    np.savez_compressed(os.path.join(HERE, "rnnt_loss.npz"), **out)
    print("rnnt done")


def _bmuf_ref_worker(rank, world, port, q):
    """one gloo rank running the REFERENCE trainer/bmuf.py:BmufTrainer (repairs: backend nccl -> gloo, .cuda() -> CPU)"""
    import torch.distributed as dist
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    ref_shim.install()
    real_init = dist.init_process_group
    dist.init_process_group = lambda backend=None, **kw: real_init(backend="gloo", **kw)
    torch.Tensor.cuda = lambda self, *a, **k: self
    from trainer.bmuf import BmufTrainer, SUCCESS
    torch.manual_seed(100 + rank)
    model = torch.nn.Sequential(torch.nn.Linear(7, 5), torch.nn.Linear(5, 3))
    tr = BmufTrainer(0, rank, world, model, 0.9, 1.0)
    out = [tr.param.clone().numpy()]
    for it in range(3):
        g = torch.Generator().manual_seed(7 * it + rank)
        with torch.no_grad():
            vec = torch.nn.utils.parameters_to_vector(model.parameters())
            vec = vec + 0.01 * torch.randn(vec.shape, generator=g)
            torch.nn.utils.vector_to_parameters(vec, model.parameters())
        assert tr.update_and_sync() == SUCCESS
        out.append(tr.param.clone().numpy())
    q.put((rank, np.stack(out), torch.nn.utils.parameters_to_vector(model.parameters()).detach().numpy()))
    dist.destroy_process_group()


def golden_bmuf():
    """Reference BmufTrainer on 2 gloo ranks: the parameter vector after the initial broadcast and after each of three
    block syncs with rank-specific local updates (same seeds as tests/test_bmuf_gloo_cpu.py)."""
    import torch.multiprocessing as mp
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29400 + (os.getpid() % 500)
    procs = [ctx.Process(target=_bmuf_ref_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = sorted([q.get(timeout=120) for _ in procs], key=lambda r: r[0])
    for p in procs:
        p.join(30)
    # every rank ends each sync with the same parameters (broadcast), and the model holds them
    assert np.array_equal(res[0][1], res[1][1]) and np.array_equal(res[0][2], res[0][1][-1]) and np.array_equal(res[1][2], res[0][1][-1])
    np.savez_compressed(os.path.join(HERE, "bmuf_2rank.npz"), params=res[0][1])
    print("bmuf: params", res[0][1].shape, "first", res[0][1][:, 0])


def golden_mbr():
    """MBR training batch: EXECUTES the reference's own loop body (trainer/train_transducer_mbr_bmuf_otfaug.py, from the
    '#nbest genereation' comment to 'out.backward(mbr_grad)') on CPU.  The script cannot be imported (it runs its
    training loop inside a function that needs PyKaldi loaders, NCCL and CUDA), so the loop-body source text is read from
    /root/reference at generation time, dedented and exec'd in a namespace that supplies what the surrounding function
    would have: the reference Net and TransducerDecoder, an SGD optimiser, the batch, and three environment repairs --
    ``.cuda()`` / ``torch.cuda.*Tensor`` mapped to CPU, ``editdistance.eval`` (absent package) = plain Levenshtein,
    ``RNNTLoss.apply`` (warp_rnnt, CUDA only) = torchaudio's rnnt_loss(reduction='none').  No reference arithmetic is
    altered.  Stores the inputs, the N-best list, both losses and a fingerprint of EVERY parameter gradient."""
    import types
    import textwrap
    import torchaudio
    from torch.autograd import Variable
    ref_shim.load_beam_module()
    from decoder.transducer_decoder import TransducerDecoder
    import decoder.beam_transducer as bt
    from fixture_utils import decode_fixture_reinit, grad_fingerprint
    V, beam = 40, 4
    m = build_ref_model(V)
    disable_dropout(m)
    decode_fixture_reinit(m)
    d = np.load(os.path.join(HERE, "decode_small.npz"))
    x = torch.from_numpy(d["x"])
    tl = torch.from_numpy(d["tlens"]).int()
    g = torch.Generator().manual_seed(11)
    ul = torch.tensor([5, 3, 4], dtype=torch.int32)
    target = torch.full((3, 5), V, dtype=torch.long)
    for i in range(3):
        target[i, :ul[i]] = torch.randint(1, V, (int(ul[i]),), generator=g)
    args = types.SimpleNamespace(local_rank=0, rnnt_scale=0.5, sm_scale=0.8, blk=0, padding_idx=V, grad_clip=0.0,
                                 las_rescorer=None, las_rescorer_bw=None, bilas_rescorer=None, nonblk_reward=0.0)
    dec = TransducerDecoder(m, 3, beam, n_best=beam, blk=0, global_scorer=bt.GlobalScorer(), sm_scale=args.sm_scale,
                            cuda=False, beam_prune=False, args=args)

    def lev(a, b):
        dd = np.zeros((len(a) + 1, len(b) + 1), dtype=np.int64)
        dd[:, 0] = np.arange(len(a) + 1)
        dd[0, :] = np.arange(len(b) + 1)
        for i in range(1, len(a) + 1):
            for j in range(1, len(b) + 1):
                dd[i, j] = min(dd[i - 1, j] + 1, dd[i, j - 1] + 1, dd[i - 1, j - 1] + (a[i - 1] != b[j - 1]))
        return int(dd[len(a), len(b)])

    def transducer_loss(log_probs, labels, frame_lens, label_lens):
        return torchaudio.functional.rnnt_loss(log_probs, labels.int(), frame_lens.int(), label_lens.int(), blank=0,
                                               reduction="none", fused_log_softmax=False)

    src = open(ref_shim.REF + "/trainer/train_transducer_mbr_bmuf_otfaug.py").read().split("\n")
    i0 = next(i for i, l in enumerate(src) if "#nbest genereation" in l)
    i1 = next(i for i, l in enumerate(src) if "out.backward(mbr_grad)" in l)
    body = textwrap.dedent("\n".join(src[i0:i1 + 1]))
    ns = dict(model=m, trans_decoder=dec, data_batch=x.clone(), len_batch=tl.clone(), ali_lens=ul.clone(),
              target_batch=target.clone(), optimizer=torch.optim.SGD(m.parameters(), 1e-3), spec_augmentor=None, args=args,
              transducer_loss=transducer_loss, beam_size=beam, F=F, torch=torch, Variable=Variable,
              editdistance=types.SimpleNamespace(eval=lev))
    saved = (torch.Tensor.cuda, torch.cuda.FloatTensor, torch.cuda.LongTensor)
    torch.Tensor.cuda = lambda self, *a, **k: self
    torch.cuda.FloatTensor, torch.cuda.LongTensor = torch.FloatTensor, torch.LongTensor
    try:
        exec(compile(body, "reference_mbr_loop_body", "exec"), ns)
    finally:
        torch.Tensor.cuda, torch.cuda.FloatTensor, torch.cuda.LongTensor = saved
    hyps = [[[int(t) for t in h] for h in row] for row in ns["hyps"]]
    scores = [[float(sc) for sc in row] for row in ns["scores"]]
    L = max(len(h) for row in hyps for h in row)
    hyp_arr = np.full((3, beam, L), -2, np.int64)
    for i in range(3):
        for j in range(beam):
            hyp_arr[i, j, :len(hyps[i][j])] = hyps[i][j]
    out = dict(target=target.numpy(), ulens=ul.numpy(), hyps=hyp_arr, scores=np.array(scores, np.float64),
               mbr_loss=np.array(float(ns["mbr_loss"])), rnnt_loss=np.array(float(ns["rnnt_loss"])),
               rnnt_scale=np.array(args.rnnt_scale), sm_scale=np.array(args.sm_scale))
    n_zero = 0
    for k, prm in m.named_parameters():
        gr = prm.grad if prm.grad is not None else torch.zeros_like(prm)
        n_zero += int(prm.grad is None)
        out["g:" + k] = grad_fingerprint(gr)
    np.savez_compressed(os.path.join(HERE, "mbr_small.npz"), **out)
    print("mbr: loss %.6f rnnt %.6f, %d params (%d without grad), hyps lens %s" %
          (out["mbr_loss"], out["rnnt_loss"], len([k for k in out if k.startswith("g:")]), n_zero, [len(h) for h in hyps[0]]))


if __name__ == "__main__":
    which = sys.argv[1:] or ["rnnt", "frontend", "specaug", "encoder", "model", "decode", "mbr", "bmuf"]
    table = dict(rnnt=golden_rnnt, frontend=golden_frontend, specaug=golden_specaug,
                 encoder=golden_encoder_eval, model=golden_model, decode=golden_decode, mbr=golden_mbr, bmuf=golden_bmuf,
                 model_full=golden_model_full, decode_big=golden_decode_big, decode_fst=golden_decode_fst, model_xf=golden_model_xf,
                 decode_xf=golden_decode_xf)
    for w in which:
        table[w]()
