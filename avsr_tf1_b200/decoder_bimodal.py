"""WLAS dual-attention decoder - drop-in for reference avsr/decoder_bimodal.py
(Seq2SeqBimodalDecoder :10-470): two mechanisms (video memory first, then audio)
inside one AttentionWrapper, 2x256 concatenated attention, initial state = one shared
no-bias Dense over concat(video_final, audio_final) (:125-166, :480-492)."""
from __future__ import annotations

from . import ops
from .decoder_unimodal import Seq2SeqUnimodalDecoder


class Seq2SeqBimodalDecoder(Seq2SeqUnimodalDecoder):
    LENGTH_PENALTY = 0.5  # decoder_bimodal.py:366

    def __init__(self, video_depth, audio_depth, video_state_depth, audio_state_depth, mode, hparams, ctx=None):
        """video_depth / audio_depth None: that stream is missing (decoder_bimodal.py:127-142) - its state is a zero
        tuple of width *_state_depth in the shared projection and it gets no attention mechanism (:179-225)."""
        self._state_depths = (int(video_state_depth), int(audio_state_depth))
        self._present = tuple(k for k, d in enumerate((video_depth, audio_depth)) if d is not None)
        if not self._present:
            raise Exception('labels are None')
        depths = [d for d in (video_depth, audio_depth) if d is not None]
        super(Seq2SeqBimodalDecoder, self).__init__(depths, mode, hparams, ctx=ctx)

    def _attention_types(self):
        at = self._hparams.attention_type
        # decoder_bimodal.py:196-223: video mechanisms use attention_type[0], audio attention_type[1]
        return [at[k][0] for k in self._present]

    def _mem_layer_names(self):
        return ['Decoder/memory_layer/kernel', 'Decoder/memory_layer_1/kernel'][:len(self._present)]

    def _extra_decls(self):
        self._Wp = self._ctx.declare('Decoder/state_projection/kernel', (sum(self._state_depths), self._H), 'glorot')

    def _initial_state_fwd(self, encoder_states):
        ctx = self._ctx
        if len(self._present) == 1:  # zero state for the missing stream
            (c1, h1), = encoder_states
            z = ops.zeros(c1.shape[0], self._state_depths[1 - self._present[0]])
            encoder_states = [(c1, h1), (z, z)] if self._present[0] == 0 else [(z, z), (c1, h1)]
        (cv, hv), (ca, ha) = encoder_states
        B, Hv = cv.shape
        Ha = ca.shape[1]
        self._cat_c, self._cat_h = ops.empty(B, Hv + Ha), ops.empty(B, Hv + Ha)
        self._cat_c[:, :Hv].copy_(cv); self._cat_c[:, Hv:].copy_(ca)
        self._cat_h[:, :Hv].copy_(hv); self._cat_h[:, Hv:].copy_(ha)
        c0, h0 = ops.empty(B, self._H), ops.empty(B, self._H)
        ops.gemm(self._cat_c, ctx.w(self._Wp), c0)
        ops.gemm(self._cat_h, ctx.w(self._Wp), h0)
        return (c0, h0)

    def _initial_state_bwd(self, dinit):
        ctx = self._ctx
        dc0, dh0 = dinit
        B = dc0.shape[0]
        Hv, Ha = self._state_depths
        ops.gemm(self._cat_c, dc0, ctx.g(self._Wp), ta=True, beta=1.0)
        ops.gemm(self._cat_h, dh0, ctx.g(self._Wp), ta=True, beta=1.0)
        dcc, dch = ops.empty(B, Hv + Ha), ops.empty(B, Hv + Ha)
        ops.gemm(dc0, ctx.w(self._Wp), dcc, tb=True)
        ops.gemm(dh0, ctx.w(self._Wp), dch, tb=True)
        both = [(dcc[:, :Hv].contiguous(), dch[:, :Hv].contiguous()),
                (dcc[:, Hv:].contiguous(), dch[:, Hv:].contiguous())]
        return [both[k] for k in self._present]
