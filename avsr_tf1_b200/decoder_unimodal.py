"""LAS-style attention decoder - drop-in for reference avsr/decoder_unimodal.py
(Seq2SeqUnimodalDecoder :9-367).  The bimodal (WLAS) decoder reuses the same
machinery with two memories (decoder_bimodal.py)."""
from __future__ import annotations

import numpy as np
import torch

from . import ops
from .attention import add_attention
from .cells import build_rnn_layers
from .layers import BuildContext


class BeamSearchOutput(object):
    """seq2seq.BeamSearchDecoderOutput fields the caller reads (avsr.py:366, 476-478)."""

    def __init__(self, scores, predicted_ids, parent_ids):
        self.scores, self.predicted_ids, self.parent_ids = scores, predicted_ids, parent_ids


def gather_tree(step_ids, parent_ids, max_len, end_token):
    """tf.contrib.seq2seq gather_tree (integer back-trace, host side).
    step_ids / parent_ids [T,B,W]; max_len [B].  Returns [T,B,W]."""
    T, B, W = step_ids.shape
    out = np.full((T, B, W), end_token, step_ids.dtype)
    for b in range(B):
        ml = int(min(max_len[b], T))
        if ml <= 0:
            continue
        for w in range(W):
            parent = parent_ids[ml - 1, b, w]
            out[ml - 1, b, w] = step_ids[ml - 1, b, w]
            for lvl in range(ml - 2, -1, -1):
                out[lvl, b, w] = step_ids[lvl, b, parent]
                parent = parent_ids[lvl, b, parent]
            hit = np.nonzero(out[:ml, b, w] == end_token)[0]
            if hit.size:
                out[hit[0]:ml, b, w] = end_token
    return out


class Seq2SeqUnimodalDecoder(object):
    LENGTH_PENALTY = 0.6  # decoder_unimodal.py:256

    def __init__(self, encoder_output_depths, mode, hparams, ctx: BuildContext = None, state_depth=None):
        """encoder_output_depths: feature size of each attended memory (1 entry here, 2 for WLAS)."""
        self._mode, self._hparams, self._ctx = mode, hparams, ctx
        reverse_dict = {v: k for k, v in hparams.unit_dict.items()}
        self._GO_ID = reverse_dict['GO']
        self._EOS_ID = reverse_dict['EOS']
        self._sampling_probability_outputs = hparams.sampling_probability_outputs
        self._vocab_size = len(hparams.unit_dict) - 1  # decoder_unimodal.py:49
        self._mem_depths = list(encoder_output_depths)
        self._init_embedding()
        self._init_decoder()
        # ScheduledEmbeddingTrainingHelper(sampling_probability) (decoder_unimodal.py:304-309): TF's Philox draws are
        # not reproducible; the library's counter-based generator stands in (streams: +0 Bernoulli select, +1 draw)
        self._ss_thr, self._ss_stream = 0, None
        if mode == 'train' and self._sampling_probability_outputs > 0.0:
            p = float(self._sampling_probability_outputs)
            self._ss_thr = max(1, min(int(p * 4294967296.0), 4294967295))
            self._ss_stream = ctx.new_stream(2)
            ctx.streams['Decoder/sampling'] = self._ss_stream
        self.sample_ids = None  # [T,B] int32 device: ids drawn by the helper (-1 = ground truth kept), last train step
        self.inference_predicted_ids = None
        self.beam_search_output = None
        self.attention_alignment = self.attention_summary = None

    def _init_embedding(self):
        hp, ctx = self._hparams, self._ctx
        if hp.embedding_size <= 0:
            # decoder_unimodal.py:75-76: one-hot inputs - the "embedding matrix" is the constant tf.eye(vocab_size): no
            # variable, nothing to train or to save
            self._E = self._vocab_size
            self._embedding = None
            self._eye = None
            return
        self._E = hp.embedding_size
        self._embedding = ctx.declare('embeddings/embedding_matrix', (self._vocab_size, self._E), 'embedding')

    def _table(self):
        """The lookup table as a product operand: the trained embedding matrix, or the identity for one-hot inputs."""
        if self._embedding is not None:
            return self._ctx.w(self._embedding)
        if self._eye is None:
            self._eye = torch.eye(self._vocab_size, dtype=torch.float32, device='cuda')
        return self._eye

    def _attention_types(self):
        return [self._hparams.attention_type[1][0]]

    def _mem_layer_names(self):
        return ['Decoder/memory_layer/kernel']

    def _init_decoder(self):
        hp, ctx = self._hparams, self._ctx
        if len(hp.decoder_units_per_layer) != 1:
            raise NotImplementedError('multi-layer decoders are not used by any reference config')
        if hp.decoding_algorithm not in ('greedy', 'beam_search'):
            raise Exception('The only supported algorithms are `greedy` and `beam_search`')
        cell = build_rnn_layers(cell_type=hp.cell_type, num_units_per_layer=hp.decoder_units_per_layer,
                                use_dropout=hp.use_dropout, dropout_probability=hp.decoder_dropout_probability,
                                mode=self._mode)
        self._H = cell.num_units
        self._extra_decls()
        if hp.enable_attention is True:
            self._cell = add_attention(cell, attention_types=self._attention_types(),
                                       num_units=hp.decoder_units_per_layer[-1], memory_depths=self._mem_depths, ctx=ctx,
                                       wrap_prefix='Decoder/decoder/attention_wrapper',
                                       mem_layer_names=self._mem_layer_names(), in_dim=self._E)
        else:
            # decoder_unimodal.py:319-327 / decoder_bimodal.py:261-263: the bare (dropout-wrapped) cell under BasicDecoder,
            # started from the decoder initial state; the encoder outputs are not attended to (`Decoder/decoder/lstm_cell`)
            from .layers import AttnLSTMOp
            self._cell = AttnLSTMOp(ctx, 'Decoder/decoder', self._E, cell.num_units, [],
                                    drop=ctx.drop_state(cell, 'Decoder/decoder'))
        self._Wd = ctx.declare('Decoder/decoder/my_dense/kernel', (self._cell.out_dim, self._vocab_size), 'glorot')
        self._bd = ctx.declare('Decoder/decoder/my_dense/bias', (self._vocab_size,), 'zeros')

    def _extra_decls(self):
        pass

    # ---- initial state (decoder_unimodal.py:126-157): copy of the encoder's last-layer state
    def _initial_state_fwd(self, encoder_states):
        c, h = encoder_states[0]
        if c.shape[1] != self._H:
            raise ValueError('decoder units must equal the encoder state size when the state is copied')
        return (c, h)

    def _initial_state_bwd(self, dinit):
        return [dinit]

    # ---- training ------------------------------------------------------------
    def forward_train(self, memories, encoder_states, dec_in_ids, labels, labels_len, T, inv_denom, loss_sum):
        """memories [(values [Tm,B,Dm], len)]; dec_in_ids [T,B] int32 (GO-prefixed labels, frame-major);
        labels [B,L] int32; accumulates the cross-entropy sum into loss_sum[0]."""
        ctx = self._ctx
        B = labels.shape[0]
        init = self._initial_state_fwd(encoder_states)
        self._n_mem = len(memories)
        if not self._cell.mechs:
            memories = []
        O = self._cell.out_dim
        self._logits = ops.empty(T, B, self._vocab_size)
        if self._ss_thr:
            out = self._forward_train_sampled(memories, init, dec_in_ids, labels_len, T, B)
        else:
            self._ids = dec_in_ids.reshape(-1)
            x = ops.empty(T, B, self._E)
            ops.embedding_fwd(self._table(), self._ids, x)
            out = self._cell.forward(x, labels_len, memories=memories, init=init)
            ops.gemm(out.view(T * B, O), ctx.p(self._Wd), self._logits.view(T * B, self._vocab_size),
                     bias=ctx.p(self._bd))
        self._out = out
        self._dlogits = torch.empty_like(self._logits)
        if self._hparams.loss_fun is not None:  # devel.py focal_loss / mc_loss (seq2seq.py:156-163)
            ops.seq_loss_devel(self._logits, labels, labels_len, inv_denom, loss_sum, self._dlogits, self._hparams.loss_fun)
        else:
            ops.seq_loss(self._logits, labels, labels_len, inv_denom, loss_sum, self._dlogits,
                         label_smoothing=float(self._hparams.label_smoothing))
        return self._logits

    def _forward_train_sampled(self, memories, init, dec_in_ids, labels_len, T, B):
        """Scheduled sampling: step t's logits decide (per row, with probability p) the input of step t + 1, so the
        recurrence advances one step at a time (AvsrRnnSeq step ranges) with the output layer and the draw in between.
        The ids actually fed are kept for the embedding gradient; no gradient flows through the draws."""
        ctx, cell = self._ctx, self._cell
        V = self._vocab_size
        used = dec_in_ids.clone()  # [T,B]; rows 1.. are overwritten where the helper draws
        self.sample_ids = torch.full((T, B), -1, dtype=torch.int32, device='cuda')
        table = self._table()
        # the persistent kernel draws inside the recurrence when it covers this cell (one launch instead of ~11 per step)
        out = cell.forward_sampled(table, dec_in_ids, labels_len, memories, init, ctx.p(self._Wd), ctx.p(self._bd),
                                   self._ss_stream, self._ss_thr, used, self.sample_ids)
        if out is not None:
            O = cell.out_dim
            ops.gemm(out.view(T * B, O), ctx.p(self._Wd), self._logits.view(T * B, V), bias=ctx.p(self._bd))
            self._ids = used.reshape(-1)
            self.decoder_input_ids = used
            return out
        cell.begin_stepwise(T, B, labels_len, memories, init)
        for t in range(T):
            x_t = ops.empty(B, self._E)
            ops.embedding_fwd(table, used[t], x_t)
            out_t = cell.stepwise_step(t, x_t)
            ops.gemm(out_t, ctx.p(self._Wd), self._logits[t], bias=ctx.p(self._bd))
            if t + 1 < T:
                ops.sched_sample(self._logits[t], ctx.rng, self._ss_stream, t, self._ss_thr, dec_in_ids[t + 1],
                                 used[t + 1], self.sample_ids[t])
        self._ids = used.reshape(-1)
        self.decoder_input_ids = used  # [T,B]: what the decoder was fed (parity probe)
        return cell.end_stepwise()

    def backward_train(self):
        """Returns ([dmemory...], [(dc, dh) per encoder state])."""
        ctx = self._ctx
        T, B, V = self._logits.shape
        O = self._cell.out_dim
        dl2 = self._dlogits.view(T * B, V)
        ops.gemm(self._out.view(T * B, O), dl2, ctx.g(self._Wd), ta=True, beta=1.0)
        ops.colsum(dl2, ctx.g(self._bd))
        dout = ops.empty(T, B, O)
        ops.gemm(dl2, ctx.p(self._Wd), dout.view(T * B, O), tb=True)
        dx, dmem, dinit = self._cell.backward(dout, None, need_dx=True, want_init_grad=True)
        if self._embedding is not None:
            ops.embedding_bwd(dx.view(T * B, self._E), self._ids, ctx.g(self._embedding))
        if not self._cell.mechs:  # nothing attends to the encoder outputs
            dmem = [None] * self._n_mem
        return dmem, self._initial_state_bwd(dinit)

    # ---- inference -------------------------------------------------------------
    def _logits_step(self, out):
        B = out.shape[0]
        logits = ops.empty(B, self._vocab_size)
        ops.gemm(out, self._ctx.p(self._Wd), logits, bias=self._ctx.p(self._bd))
        return logits

    def decode_greedy(self, memories, encoder_states):
        """GreedyEmbeddingHelper + dynamic_decode(impute_finished=True) (decoder_unimodal.py:176-220)."""
        ctx, hp = self._ctx, self._hparams
        B = memories[0][0].shape[1]
        init = self._initial_state_fwd(encoder_states)
        if not self._cell.mechs:
            memories = []
        bufs = self._cell.prepare_memories(memories)
        state = self._cell.initial_state(B, init)
        ids = torch.full((B,), self._GO_ID, dtype=torch.int32, device='cuda')
        finished = torch.zeros(B, dtype=torch.int32, device='cuda')
        active = torch.ones(B, dtype=torch.int32, device='cuda')
        samples = []
        history = [[] for _ in bufs] if (hp.write_attention_alignment and bufs) else None
        for _ in range(hp.max_label_length):
            x = ops.empty(1, B, self._E)
            ops.embedding_fwd(self._table(), ids, x)
            torch.sub(1, finished, out=active)  # finished rows carry their state (impute_finished)
            out, state = self._cell.step(x, active, bufs, state)
            if history is not None:  # alignment_history (attention.py:178)
                for k, mb in enumerate(bufs):
                    history[k].append(mb.align[0].clone())
            logits = self._logits_step(out)
            sample = torch.empty(B, dtype=torch.int32, device='cuda')
            nxt = torch.empty(B, dtype=torch.int32, device='cuda')
            ops.greedy_pick(logits, self._EOS_ID, finished, sample, nxt)
            samples.append(sample)
            ids = nxt
            if bool(finished.all().item()):
                break
        self.inference_predicted_ids = torch.stack(samples, dim=1).cpu().numpy().astype(np.int32)
        if history is not None:
            # _create_attention_alignments_summary (decoder_unimodal.py:273-290, decoder_bimodal.py:447-467):
            # [T, B, Tm] -> [B, Tm, T, 1]; one per mechanism for the bimodal decoder (video first)
            al = [torch.stack(h, 0).permute(1, 2, 0).unsqueeze(-1).cpu().numpy() for h in history]
            self.attention_alignment = al[0] if len(al) == 1 else al
            self.attention_summary = 1.0 - al[0] if len(al) == 1 else [1.0 - a for a in al]
        return self.inference_predicted_ids

    def decode_beam(self, memories, encoder_states):
        """BeamSearchDecoder(beam_width, length_penalty_weight) + gather_tree
        (decoder_unimodal.py:222-271).  Returns beam 0 ids [B, T]."""
        ctx, hp = self._ctx, self._hparams
        # Alignment images under beam search: the reference's own branch (decoder_unimodal.py:277-280) subscripts the
        # TensorArray `cell_state.alignment_history` of a state that BeamSearchDecoder never reorders and cannot run in
        # TF 1.13.  Provided here as what that branch is after: the alignments of the WINNING hypothesis, traced back along
        # the beam parents like its ids (gather_tree), in the layout of the greedy images [B, Tm, T, 1].
        W = hp.beam_width
        B = memories[0][0].shape[1]
        init = self._initial_state_fwd(encoder_states)
        if not self._cell.mechs:
            memories = []
        tiled = [(m[0].repeat_interleave(W, dim=1).contiguous(), m[1].repeat_interleave(W).contiguous(),
                  m[2].repeat_interleave(W, dim=1).contiguous() if len(m) > 2 and m[2] is not None else None)
                 for m in memories]  # seq2seq.tile_batch (attention.py:101-106)
        init = (init[0].repeat_interleave(W, dim=0).contiguous(), init[1].repeat_interleave(W, dim=0).contiguous())
        bufs = self._cell.prepare_memories(tiled)
        c, S = self._cell.initial_state(B * W, init)
        log_probs = torch.full((B, W), float('-inf'), device='cuda')
        log_probs[:, 0] = 0.0
        finished = torch.zeros((B, W), dtype=torch.int32, device='cuda')
        lengths = torch.zeros((B, W), dtype=torch.int32, device='cuda')
        ids = torch.full((B * W,), self._GO_ID, dtype=torch.int32, device='cuda')
        active = torch.ones(B * W, dtype=torch.int32, device='cuda')
        base = (torch.arange(B, device='cuda', dtype=torch.int32) * W).view(B, 1)
        words, parents, scores = [], [], []
        history = [[] for _ in bufs] if (hp.write_attention_alignment and bufs) else None
        for _ in range(hp.max_label_length):
            x = ops.empty(1, B * W, self._E)
            ops.embedding_fwd(self._table(), ids, x)
            out, (c, S) = self._cell.step(x, active, bufs, (c, S))
            if history is not None:  # alignments of every live hypothesis at this step, by beam slot BEFORE the re-ranking
                for k, mb in enumerate(bufs):
                    history[k].append(mb.align[0].clone())
            logits = self._logits_step(out)
            word = torch.empty((B, W), dtype=torch.int32, device='cuda')
            parent = torch.empty((B, W), dtype=torch.int32, device='cuda')
            score = torch.empty((B, W), device='cuda')
            ops.beam_step(logits, B, W, self._EOS_ID, self.LENGTH_PENALTY, log_probs, finished, lengths, word, parent,
                          score)
            flat = (base + parent).reshape(-1).contiguous()
            c2, S2 = torch.empty_like(c), torch.empty_like(S)
            ops.gather_rows(c, flat, c2)
            ops.gather_rows(S.contiguous(), flat, S2)
            c, S = c2, S2
            ids = word.reshape(-1)
            words.append(word)
            parents.append(parent)
            scores.append(score)
            if bool(finished.all().item()):
                break
        step_ids = torch.stack(words, 0).cpu().numpy()
        parent_ids = torch.stack(parents, 0).cpu().numpy()
        max_len = lengths.max(dim=1).values.cpu().numpy()
        pred = gather_tree(step_ids, parent_ids, max_len, self._EOS_ID)  # [T,B,W]
        self.inference_predicted_beam = pred.transpose(1, 0, 2).astype(np.int32)
        self.inference_predicted_ids = self.inference_predicted_beam[:, :, 0]
        self.beam_search_output = BeamSearchOutput(
            scores=torch.stack(scores, 0).cpu().numpy().transpose(1, 0, 2),
            predicted_ids=step_ids.transpose(1, 0, 2).astype(np.int32),
            parent_ids=parent_ids.transpose(1, 0, 2).astype(np.int32))
        if history is not None:
            al = []
            for h in history:
                A = torch.stack(h, 0).cpu().numpy()           # [T, B*W, Tm]
                T_, Tm = A.shape[0], A.shape[2]
                A = A.reshape(T_, B, W, Tm)
                out_al = np.zeros((B, Tm, T_, 1), A.dtype)
                for b in range(B):
                    ml = int(min(max_len[b], T_))
                    slot = 0                                   # winning hypothesis = beam 0 after the last re-ranking
                    for lvl in range(ml - 1, -1, -1):
                        slot = int(parent_ids[lvl, b, slot])   # the slot it occupied when step lvl ran
                        out_al[b, :, lvl, 0] = A[lvl, b, slot]
                al.append(out_al)
            self.attention_alignment = al[0] if len(al) == 1 else al
            self.attention_summary = 1.0 - al[0] if len(al) == 1 else [1.0 - a for a in al]
        return self.inference_predicted_ids

    def get_predictions(self):
        return self.inference_predicted_ids
