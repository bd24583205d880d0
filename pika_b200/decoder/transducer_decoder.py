"""Transducer beam-search decoder -- drop-in for decoder/transducer_decoder.py (reference).

Same constructor and ``decode_batch(x, x_len, max_len) -> (ret, enc_out)`` contract
(``ret = {"predictions": B x n_best lists of 0-d int64 tensors (full alignment incl. blanks, trailing EOS
stripped), "scores": B x n_best 0-d f32 tensors}``).  All per-step work runs on the GPU for the whole batch:
encoder-frame gather, masked LSTM step, factored joint, log-softmax, and one ``pk_beam_advance`` launch that
performs every utterance's score add / EOS + duplicate kill / top-k / finish rule / hypothesis update
(decoder/beam_transducer.py:82-187).  The host reads one "utterances not done" counter per step and walks the
back-pointers once at the end (decoder/transducer_decoder.py:204-217, decoder/beam_transducer.py:196-243).

Like the reference, every utterance keeps advancing until ALL utterances of the batch are done, so late
finishes can still enter an utterance's n-best list.
"""
import ctypes

import numpy as np
import torch

from .. import engine
from .. import kernels as K
from .._lib import check, lib


class TransducerDecoder():
    def __init__(self, model, batch_size, beam_size, n_best=1, blk=0, global_scorer=None, sm_scale=1.0, lm=None, lm_scale=1.0,
                 lm_scorer=None, lm_scorer_scale=1.0, cuda=False, beam_prune=True, args=None):
        self.model, self.batch_size, self.beam_size, self.n_best, self.blk = model, batch_size, beam_size, n_best, blk
        self.global_scorer, self.sm_scale, self.cuda, self.beam_prune, self.args = global_scorer, sm_scale, cuda, beam_prune, args
        if lm is not None or lm_scorer is not None:
            raise NotImplementedError("pika_b200: LM shallow fusion is outside the hot path (SURVEY.md section 8f)")
        for name in ("las_rescorer", "las_rescorer_bw", "bilas_rescorer"):
            if args is not None and getattr(args, name, None) is not None:
                raise NotImplementedError("pika_b200: LAS rescoring is outside the hot path")
        if model.decoder_type != "rnn":
            raise NotImplementedError("pika_b200: only the LSTM prediction net is supported")

    @torch.no_grad()
    def decode_batch(self, x, x_len, max_len=None, enc_out=None):
        """``enc_out`` (extension): encoder outputs [B, T', H] computed by the caller (the MBR step shares them with the
        training forward); ``x`` is then ignored."""
        m, Kb, V, blk = self.model, self.beam_size, self.model.fc2.weight.shape[0], self.blk
        dev = x.device if enc_out is None else enc_out.device
        assert dev.type == "cuda", "pika_b200 decodes on the GPU (there is no CPU fallback)"
        if enc_out is None:
            enc = engine.encoder_forward_act(m.encoder, x).contiguous()          # [B, T', H]
        else:
            enc = engine._to_act(enc_out)
        B, Tenc, H = enc.shape
        rows = B * Kb
        adt = enc.dtype
        lstm, L = m.decoder, m.decoder.num_layers
        E = m.embed.weight.shape[1]
        ldx = (E + 7) // 8 * 8
        P, st = K._P, K._stream
        i32 = lambda *s, fill=0: torch.full(s, fill, dtype=torch.int32, device=dev)
        nf = torch.as_tensor(np.asarray([int(v) for v in x_len]), dtype=torch.int32).to(dev)
        ml_list = [int(max_len[i]) if max_len[i] else 10000 for i in range(B)]
        ml = torch.tensor(ml_list, dtype=torch.int32, device=dev)
        S = max(ml_list) + 2
        cap = S * Kb
        next_ys, prev_ks = i32(S + 1, B, Kb, fill=blk), i32(S, B, Kb)
        hyp_tok, hyp_len = i32(2, B, Kb, S + 1), i32(2, B, Kb)
        fin_score = torch.zeros(B, cap, dtype=torch.float32, device=dev)
        fin_step, fin_k, fin_count, eos_top, done = i32(B, cap), i32(B, cap), i32(B), i32(B), i32(B)
        not_done = i32(1, fill=B)
        scores = torch.zeros(B, Kb, dtype=torch.float32, device=dev)
        t_idx, t_alt = i32(rows, fill=-1), i32(rows)
        h = torch.zeros(L, rows, H, dtype=adt, device=dev)
        c = torch.zeros(L, rows, H, dtype=torch.float32, device=dev)
        h_alt, c_alt = torch.empty_like(h), torch.empty_like(c)
        enc_hid = torch.empty(rows, H, dtype=adt, device=dev)
        x_emb = torch.empty(rows, ldx, dtype=adt, device=dev)
        gates = torch.empty(rows, 4 * H, dtype=torch.float32, device=dev)
        pre = torch.empty(rows, 2 * H, dtype=torch.float32, device=dev)
        hj = torch.empty(rows, H, dtype=adt, device=dev)
        ldv = (V + 3) // 4 * 4
        logits = torch.zeros(rows, ldv, dtype=torch.float32, device=dev)
        wp = torch.empty(rows, V, dtype=torch.float32, device=dev)
        # staged weights
        w_ih = [engine.stage_weight(getattr(lstm, "weight_ih_l%d" % l), cols_pad=ldx if l == 0 and ldx != E else None) for l in range(L)]
        w_hh = [engine.stage_weight(getattr(lstm, "weight_hh_l%d" % l)) for l in range(L)]
        bsum = []
        for l in range(L):
            bs = torch.empty(4 * H, dtype=torch.float32, device=dev)
            K.add(getattr(lstm, "bias_ih_l%d" % l).detach(), getattr(lstm, "bias_hh_l%d" % l).detach(), bs)
            bsum.append(bs)
        wx = engine.stage_weight([m.fc1.weight, m.fc_gate.weight])
        bx = engine._cat_bias([m.fc1.bias, m.fc_gate.bias])
        w2 = engine.stage_weight(m.fc2.weight)

        def lstm_step(tok, masked):
            xin = x_emb
            for l in range(L):
                engine.gemm_parts([engine.stage_act(xin)], [w_ih[l]], gates, bias=bsum[l])
                if masked:                                                    # (zero state at initialisation: no h W_hh term)
                    engine.gemm_parts([engine.stage_act(h[l])], [w_hh[l]], gates, accumulate=True, k_splits=1)
                    check(lib.pk_beam_lstm_cell(P(gates), P(tok), blk, P(h[l]), K._dt(h), P(c[l]), rows, H, st()), "pk_beam_lstm_cell")
                else:
                    K.lstm_cell_fwd(gates, None, None, c[l], h[l], None, rows, H)
                xin = h[l]

        # (3) initial decoder state = LSTM(embed(blk)) from zeros (decoder/transducer_decoder.py:116)
        x_emb.zero_()
        x_emb[:, :E] = m.embed.weight.detach()[blk].to(adt)
        lstm_step(None, masked=False)

        step = 0
        while step < S - 1:
            tok = next_ys[step].reshape(-1)
            check(lib.pk_beam_prepare(P(tok), P(t_idx), P(enc), K._dt(enc), Tenc, H, P(enc_hid), P(m.embed.weight.detach()), E,
                                      P(x_emb), ldx, Kb, blk, rows, st()), "pk_beam_prepare")
            lstm_step(tok, masked=True)
            engine.gemm_parts([engine.stage_act(enc_hid), engine.stage_act(h[L - 1])],
                              [[p[:, :H] for p in wx], [p[:, H:] for p in wx]], pre, bias=bx)
            check(lib.pk_beam_gate(P(pre), P(hj), K._dt(hj), rows, H, st()), "pk_beam_gate")
            engine.gemm_parts([engine.stage_act(hj)], [w2], logits[:, :V], bias=m.fc2.bias.detach())
            K.log_softmax(logits, wp, V, self.sm_scale)
            check(lib.pk_beam_advance(P(wp), P(t_idx), P(nf), P(ml), P(scores), P(next_ys), P(prev_ks), P(hyp_tok), P(hyp_len),
                                      P(fin_score), P(fin_step), P(fin_k), P(fin_count), P(eos_top), P(done), P(not_done), B, Kb, V,
                                      S + 1, cap, step, blk, self.n_best, int(bool(self.beam_prune)), st()), "pk_beam_advance")
            check(lib.pk_beam_reorder(P(prev_ks[step]), P(h), P(c), P(t_idx), P(h_alt), P(c_alt), P(t_alt), K._dt(h), Kb, H, L, rows,
                                      st()), "pk_beam_reorder")
            h, h_alt, c, c_alt, t_idx, t_alt = h_alt, h, c_alt, c, t_alt, t_idx
            step += 1
            if int(not_done.item()) == 0:                                     # `while not all(b.done() for b in beam)`
                break

        # (4) extract: sort_finished + get_hyp on the host, once
        ny, pk = next_ys[:step + 1].cpu().numpy(), prev_ks[:step].cpu().numpy()
        fs, fstep, fk, fc = fin_score.cpu().numpy(), fin_step.cpu().numpy(), fin_k.cpu().numpy(), fin_count.cpu().numpy()
        ret = {"predictions": [], "scores": []}
        for b in range(B):
            n = int(fc[b])
            order = sorted(range(n), key=lambda i: -fs[b, i])                 # stable, like list.sort(key=-score)
            hyps, scs = [], []
            for i in order[:self.n_best]:
                k, toks = int(fk[b, i]), []
                for j in range(int(fstep[b, i]) - 1, -1, -1):                 # BeamMergeTransducer.get_hyp
                    toks.append(int(ny[j + 1, b, k]))
                    k = int(pk[j, b, k])
                toks = toks[::-1][:-1]                                        # strip the ending eos(-1)
                hyps.append([torch.tensor(t, dtype=torch.long) for t in toks])
                scs.append(torch.tensor(float(fs[b, i]), dtype=torch.float32))
            ret["predictions"].append(hyps)
            ret["scores"].append(scs)
        return ret, enc.float()
