"""Seq2SeqModel - drop-in for reference avsr/seq2seq.py (Seq2SeqModel :9-280).

Same constructor, same attribute surface the reference caller (avsr/avsr.py)
uses - ``train_op``, ``batch_loss``, ``global_norm``, ``global_step``, ``saver``,
``_decoder.inference_predicted_ids`` - but eager: there is no TF graph/session.
``data_sequences`` carries concrete batches (numpy or torch, batch-major like the
reference) instead of tf.data iterator nodes; ``train_op()`` runs one optimiser
step on the currently fed batch, ``feed()`` swaps the batch.

Everything numeric runs in hand-written sm_100a CUDA (csrc/) through the C ABI of
include/avsr_b200.h; torch supplies device buffers, streams and NCCL only."""
from __future__ import annotations

import math
import os
from typing import Dict

import numpy as np
import torch

from . import ops, parallel
from .decoder_bimodal import Seq2SeqBimodalDecoder
from .decoder_unimodal import Seq2SeqUnimodalDecoder
from .encoder import AttentiveEncoder, Seq2SeqEncoder
from .layers import BuildContext
from .params import ParamStore


def cosine_decay_restarts(learning_rate, global_step, first_decay_steps, t_mul=2.0, m_mul=1.0, alpha=0.0):
    """tf.train.cosine_decay_restarts (SGDR) with the TF 1.13 defaults the reference uses (seq2seq.py:266-270): cosine
    from 1 to alpha over a period, periods growing by t_mul, peak scaled by m_mul per restart."""
    completed = float(global_step) / float(first_decay_steps)
    if t_mul == 1.0:
        i_restart = math.floor(completed)
        completed -= i_restart
    else:
        i_restart = math.floor(math.log(1.0 - completed * (1.0 - t_mul)) / math.log(t_mul))
        sum_r = (1.0 - t_mul ** i_restart) / (1.0 - t_mul)
        completed = (completed - sum_r) / t_mul ** i_restart
    cosine_decayed = 0.5 * (m_mul ** i_restart) * (1.0 + math.cos(math.pi * completed))
    return learning_rate * ((1.0 - alpha) * cosine_decayed + alpha)


class _PendingScalars(object):
    """Result of Seq2SeqModel.fetch_scalars_async()."""

    def __init__(self, model, buf, event, inv_denom, au_scale):
        self._model, self._buf, self._event = model, buf, event
        self._inv_denom, self._au_scale = inv_denom, au_scale

    def result(self):
        self._event.synchronize()
        m, hp = self._model, self._model._hparams
        vals = self._buf.numpy()
        xent = float(vals[0]) * self._inv_denom
        reg = 0.5 * (hp.recurrent_l2_regularisation or 0.0) * float(vals[1]) + 0.5 * 1e-3 * float(vals[4])
        au = float(vals[3]) * self._au_scale if m._au_head else 0.0
        m.batch_loss = xent + reg + au
        m.global_norm = math.sqrt(float(vals[2]))
        return m.batch_loss, m.global_norm


class Saver(object):
    """tf.train.Saver stand-in (seq2seq.py:132-133): all global variables incl. Adam slots,
    BN moving statistics and global_step, keyed by TF variable name, in one .npz file."""

    def __init__(self, model, max_to_keep=1):
        self._model = model
        self._max_to_keep = max_to_keep  # tf.train.Saver(max_to_keep=1), seq2seq.py:133
        self._kept = []

    def save(self, sess=None, save_path=None, global_step=None):
        path = save_path if global_step is None else '%s-%d' % (save_path, global_step)
        os.makedirs(os.path.dirname(os.path.abspath(path)) or '.', exist_ok=True)
        st = self._model.store
        blob = {'var/' + k: v for k, v in st.to_numpy('p').items()}
        if st.m is not None:
            blob.update({'adam_m/' + k: v for k, v in st.to_numpy('m').items()})
            blob.update({'adam_v/' + k: v for k, v in st.to_numpy('v').items()})
        blob['global_step'] = np.asarray(self._model._global_step, np.int64)
        # written next to the target and renamed into place: a crash mid-save never leaves a truncated file as the
        # newest checkpoint (tf.train.Saver writes to a temporary name too); the old one is removed only afterwards
        tmp = path + '.tmp.npz'
        with open(tmp, 'wb') as fh:
            np.savez(fh, **blob)
            fh.flush()
            os.fsync(fh.fileno())
        os.replace(tmp, path + '.npz')
        if path not in self._kept:
            self._kept.append(path)
        while self._max_to_keep and len(self._kept) > self._max_to_keep:
            old = self._kept.pop(0) + '.npz'
            if os.path.exists(old):
                os.remove(old)
        return path

    def restore(self, sess=None, save_path=None):
        blob = np.load(save_path if save_path.endswith('.npz') else save_path + '.npz')
        st = self._model.store
        st.load_numpy({k[4:]: blob[k] for k in blob.files if k.startswith('var/')})  # also refreshes the tf32 copy
        if st.m is not None and any(k.startswith('adam_m/') for k in blob.files):
            for s in st.specs:
                if s.trainable:
                    st._view(st.m, s.name).copy_(torch.from_numpy(blob['adam_m/' + s.name]).reshape(
                        st._view(st.m, s.name).shape))
                    st._view(st.v, s.name).copy_(torch.from_numpy(blob['adam_v/' + s.name]).reshape(
                        st._view(st.v, s.name).shape))
        self._model._global_step = int(blob['global_step'])


class Seq2SeqModel(object):
    def __init__(self, data_sequences, mode, hparams, seed=2001, share_params_with=None, device='cuda'):
        self._video_data = data_sequences[0]
        self._audio_data = data_sequences[1]
        self._mode = mode
        self._hparams = hparams
        if mode not in ('train', 'evaluate'):
            raise ValueError('mode must be `train` or `evaluate`')
        if hparams.cell_type != 'lstm':
            raise Exception('cell type not supported: {}'.format(hparams.cell_type))
        self._ctx = BuildContext()
        self._ctx.world_size = parallel.world_size()
        self._ctx.allreduce = parallel.allreduce_sum_

        # Action-Unit regression head on the video encoder (train mode, seq2seq.py:188-190)
        self._au_head = bool(hparams.regress_aus) and mode == 'train' and self._video_data is not None
        self._make_encoders()
        self._make_decoder()

        if share_params_with is not None:
            self.store = share_params_with.store
        else:
            # device='cpu' builds the variable table only (tests, checkpoint tools); compute needs CUDA
            self.store = ParamStore(self._ctx.specs, device=device, with_optimizer=(mode == 'train'))
            self.store.initialize(seed, vocab=len(hparams.unit_dict) - 1)
        self._ctx.store = self.store
        self._global_step = 0
        self.batch_loss = None
        self.global_norm = None
        self.current_lr = None
        if mode == 'train':
            self._init_optimiser()
        self._init_saver()
        self._batch = None
        self._in_sets, self._in, self._meta = {}, None, None
        self._stage_sets, self._pending, self._copy_stream = {}, None, None
        self._side_stream = None
        # Independent encoder branches CAN run on two streams (video encoder next to the audio layers below the
        # cross-modal attention).  That paid off while a persistent LSTM kernel occupied 64 SMs (clusters of 8); the
        # cluster-of-4 kernels run 32 clusters on 128 SMs, two of them only time-slice the same SMs and the audio
        # branch (the critical path) loses: measured 13.6k utt/s overlapped vs 15.7k on one stream.
        self.overlap_streams = False
        self.serial_chains = False  # True: never run independent recurrent chains side by side (debugging, A/B timing)
        self._graphs = {}
        self.use_cuda_graph = False  # opt-in: train_step replays one captured graph per batch shape
        self.launches_last_step = 0
        self.h2d_bytes = 0
        # words {seed, training step} of the counter-based generator behind dropout and scheduled sampling
        # (graph-replay safe: written by _set_step_scalars outside the captured region)
        self.rng_seed = int(hparams.kwargs.get('random_seed', seed))
        if self.store.flat.is_cuda:
            self._scal_dev = torch.zeros(3, dtype=torch.float32, device='cuda')  # inv_denom, lr_t, AU scale
            self._ctx.rng = torch.zeros(2, dtype=torch.int32, device='cuda')

    # ---- construction (seq2seq.py:30-126) ---------------------------------------
    def _make_encoders(self):
        hp, ctx = self._hparams, self._ctx
        self._cnn = None
        if self._video_data is not None:
            feature_dim = None
            if hp.video_processing is not None and 'cnn' in hp.video_processing:
                # avsr.py:686-696: the lip crops [B,T,H,W,C] go through the CNN front-end, the encoder sees its features
                from .video import cnn_layers
                shape = tuple(self._video_data.inputs.shape)
                if len(shape) != 5:
                    raise Exception('`%s` needs image sequences [B,T,H,W,C], got %r' % (hp.video_processing, shape))
                self._cnn = cnn_layers(ctx, shape[2], shape[3], shape[4], hp.video_processing,
                                       hp.kwargs.get('cnn_filters', (8, 16, 32, 64)),
                                       hp.kwargs.get('cnn_dense_units', 128))
                feature_dim = self._cnn.out_dim
            self._video_encoder = Seq2SeqEncoder(
                data=self._video_data, mode=self._mode, hparams=hp,
                num_units_per_layer=hp.encoder_units_per_layer[0],
                dropout_probability=hp.video_encoder_dropout_probability, regress_aus=hp.regress_aus, ctx=ctx,
                scope='video', feature_dim=feature_dim)
            if self._cnn is not None and self._mode == 'train':
                self._video_encoder.input_gradient = True  # the CNN trains through the encoder's input
        else:
            self._video_encoder = None
        if self._audio_data is not None:
            if hp.architecture in ('unimodal', 'bimodal',):
                self._audio_encoder = Seq2SeqEncoder(
                    data=self._audio_data, mode=self._mode, hparams=hp,
                    num_units_per_layer=hp.encoder_units_per_layer[1],
                    dropout_probability=hp.audio_encoder_dropout_probability, ctx=ctx, scope='audio')
            elif hp.architecture == 'av_align':
                if self._video_encoder is None:
                    raise Exception('av_align needs a video stream')
                self._audio_encoder = AttentiveEncoder(
                    data=self._audio_data, mode=self._mode, hparams=hp,
                    num_units_per_layer=hp.encoder_units_per_layer[1],
                    attended_memory_depth=self._video_encoder.output_dim,
                    dropout_probability=hp.audio_encoder_dropout_probability, ctx=ctx, scope='audio')
            else:
                raise Exception('Unknown architecture')
        else:
            self._audio_encoder = None

    def _state_depth(self, enc):
        hp = self._hparams
        if hp.encoder_type == 'bidirectional' and not isinstance(enc, AttentiveEncoder):
            return hp.decoder_units_per_layer[0]
        return enc._num_units_per_layer[-1]

    def _make_decoder(self):
        hp, ctx = self._hparams, self._ctx
        if self._video_encoder is None and self._audio_encoder is None:
            raise Exception('labels are None')
        if hp.architecture in ('unimodal', 'av_align'):
            enc = self._audio_encoder if self._audio_encoder is not None else self._video_encoder
            self._decoder = Seq2SeqUnimodalDecoder([enc.output_dim], mode=self._mode, hparams=hp, ctx=ctx)
        elif hp.architecture == 'bimodal':
            if self._video_encoder is None or self._audio_encoder is None:
                # one stream only (decoder_bimodal.py:127-142, 179-225): the missing stream enters the shared state
                # projection as a zero state shaped like the FIRST layer's state of the present encoder
                # (`final_state[0].c`) and gets no attention mechanism
                key = 'audio' if self._video_encoder is None else 'video'
                enc = self._audio_encoder if key == 'audio' else self._video_encoder
                zero_depth = enc._num_units_per_layer[0]
                if hp.encoder_type == 'bidirectional':
                    zero_depth = self._state_depth(enc)
                depths = (self._state_depth(enc), zero_depth) if key == 'video' else (zero_depth, self._state_depth(enc))
                self._decoder = Seq2SeqBimodalDecoder(
                    enc.output_dim if key == 'video' else None, enc.output_dim if key == 'audio' else None,
                    depths[0], depths[1], mode=self._mode, hparams=hp, ctx=ctx)
                return
            self._decoder = Seq2SeqBimodalDecoder(
                self._video_encoder.output_dim, self._audio_encoder.output_dim,
                self._state_depth(self._video_encoder), self._state_depth(self._audio_encoder), mode=self._mode,
                hparams=hp, ctx=ctx)
        else:
            raise Exception('Unknown architecture')

    def _init_saver(self):
        self.saver = Saver(self)

    def _init_optimiser(self):
        hp = self._hparams
        if hp.loss_fun is not None and hp.loss_fun not in ops.LOSS_FUNS:  # focal_loss / mc_loss (seq2seq.py:156-163)
            raise ValueError('Unknown loss function {}'.format(hp.loss_fun))
        if hp.optimiser not in ops.OPTIMISERS:  # Adam, Nadam, AdamW, Momentum (seq2seq.py:195-219)
            raise Exception('Unsupported optimiser, try Adam')
        self._l2_names = [n for n in self.store.names() if 'lstm_' in n and 'bias' not in n]  # seq2seq.py:283-290
        # xent sum, L2, |g|^2, AU, CNN L2, 1 / (global token count + 1e-12) (the denominator the step actually used)
        self._loss_dev = torch.zeros(6, dtype=torch.float32, device=self.store.flat.device)

    @property
    def global_step(self):
        return self._global_step

    @property
    def n_params(self):
        return self.store.n_trainable

    # ---- batches -------------------------------------------------------------------
    def _as_tensor(self, a, dtype):
        if isinstance(a, np.ndarray):
            a = torch.from_numpy(np.ascontiguousarray(a))
        return a if a.dtype == dtype else a.to(dtype)

    def _collect(self, data_sequences):
        """Batch (batch-major like the reference; numpy / pinned host / device tensors) -> source tensors,
        host-side metadata and the shape key of the static buffers / captured graph."""
        video, audio = data_sequences
        ref = audio if audio is not None else video
        src = {}
        for key, d in (('video', video), ('audio', audio)):
            if d is None:
                continue
            # lip crops may arrive as the stored pixels (uint8): they cross PCIe as bytes and become the reference's
            # (v - 128) / 128 floats on the device (_prep)
            raw_u8 = key == 'video' and torch.is_tensor(d.inputs) and d.inputs.dtype == torch.uint8
            x = d.inputs if raw_u8 else self._as_tensor(d.inputs, torch.float32)
            if x.dim() > 3 and not (key == 'video' and self._cnn is not None):
                x = x.reshape(x.shape[0], x.shape[1], -1)  # raw lip crops fed as flat features (`features`)
            src[key] = x
            src[key + '_len'] = self._as_tensor(d.inputs_length, torch.int32)
        meta = {}
        if self._au_head:
            aus = (video.payload or {}).get('aus') if video is not None else None
            if aus is None:
                raise Exception('regress_aus=True needs the `aus` payload of the video stream (io_utils.py:45-46)')
            src['aus'] = self._as_tensor(aus, torch.float32)
            vl = video.inputs_length
            vl = vl.cpu().numpy() if torch.is_tensor(vl) else np.asarray(vl)
            meta['au_count'] = 2.0 * float(vl.sum())  # non-zero weights of tf.losses.mean_squared_error
        if ref.labels is not None:
            ll = ref.labels_length
            lab_len_host = ll.cpu().numpy() if torch.is_tensor(ll) else np.asarray(ll)
            meta['T_dec'] = int(lab_len_host.max())
            meta['n_tokens'] = float(lab_len_host.sum())
            meta['n_positions'] = float(len(lab_len_host) * int(lab_len_host.max()))
            src['labels'] = self._as_tensor(ref.labels, torch.int32)
            src['labels_len'] = self._as_tensor(ll, torch.int32)
        # data parallelism: the iterator reports the size of the GLOBAL batch this shard was cut from (payload
        # 'global_batch_size'); without it every rank is taken to hold an equal share
        gb = (ref.payload or {}).get('global_batch_size') if ref.payload is not None else None
        local_b = int(src[('audio' if audio is not None else 'video') + '_len'].shape[0])
        meta['batch_scale'] = float(gb) / local_b if gb else float(self._ctx.world_size)
        key = tuple((k, tuple(v.shape)) for k, v in sorted(src.items())) + (meta.get('T_dec', 0), meta['batch_scale'])
        meta['key'] = key
        meta['h2d_bytes'] = sum(v.numel() * v.element_size() for v in src.values() if not v.is_cuda)
        return src, meta

    MAX_INPUT_SETS = 16  # static input buffers kept per batch shape (bucketed epochs produce many shapes)

    def _static_buffers(self, key, src):
        bufs = self._in_sets.pop(key, None)
        if bufs is None:
            bufs = {k: torch.empty(v.shape, dtype=v.dtype, device='cuda') for k, v in src.items()}
            # least-recently-used shapes go first; a shape with a captured graph keeps its buffers (the graph reads them)
            for old in [k for k in self._in_sets if k not in self._graphs][:max(0, len(self._in_sets) + 1 - self.MAX_INPUT_SETS)]:
                del self._in_sets[old]
                self._stage_sets.pop(old, None)
        self._in_sets[key] = bufs  # (re)inserted last = most recently used
        return bufs

    def feed(self, data_sequences):
        """Copy one batch into the static device buffers of its shape (the layout change to frame-major
        happens on the device, _prep)."""
        src, meta = self._collect(data_sequences)
        bufs = self._static_buffers(meta['key'], src)
        for k, v in src.items():
            bufs[k].copy_(v, non_blocking=True)
        self.h2d_bytes = meta['h2d_bytes']
        self._in, self._meta = bufs, meta
        self._batch = None
        self._pending = None
        return meta

    def prefetch(self, data_sequences):
        """Input-pipeline overlap: start the host->device copy of the NEXT batch on a copy stream while the
        current step computes (two staging sets per shape).  The next train_step() without arguments picks
        it up.  Mirrors tf.data's prefetch in the reference (io_utils.py:145)."""
        if self._copy_stream is None:
            self._copy_stream = torch.cuda.Stream()
        src, meta = self._collect(data_sequences)
        key = meta['key']
        self._static_buffers(key, src)
        sets = self._stage_sets.setdefault(key, [None, None, 0])
        slot = sets[2] & 1
        sets[2] += 1
        if sets[slot] is None:
            sets[slot] = ({k: torch.empty(v.shape, dtype=v.dtype, device='cuda') for k, v in src.items()}, None)
        stage, free_ev = sets[slot]
        with torch.cuda.stream(self._copy_stream):
            if free_ev is not None:
                self._copy_stream.wait_event(free_ev)  # the compute stream has finished reading this staging set
            for k, v in src.items():
                stage[k].copy_(v, non_blocking=True)
            ready = torch.cuda.Event()
            ready.record(self._copy_stream)
        self._pending = (key, slot, meta, ready)

    def _consume_prefetch(self):
        key, slot, meta, ready = self._pending
        self._pending = None
        cur = torch.cuda.current_stream()
        cur.wait_event(ready)
        stage, _ = self._stage_sets[key][slot]
        bufs = self._in_sets[key]
        for k, v in stage.items():
            bufs[k].copy_(v, non_blocking=True)  # device-to-device into the buffers the captured graph reads
        free_ev = torch.cuda.Event()
        free_ev.record(cur)
        self._stage_sets[key][slot] = (stage, free_ev)
        self.h2d_bytes = meta['h2d_bytes']
        self._in, self._meta = bufs, meta
        self._batch = None

    def _prep(self):
        """Device-side batch preparation (capturable): GO-prefixed ids.  The features stay batch-major [B,T,F] (the
        reference's layout): the encoders change to the frame-major device layout inside their input normalisation."""
        src, meta = self._in, self._meta
        b: Dict[str, object] = {}
        for key in ('video', 'audio'):
            if key in src:
                b[key] = ops.u8_to_f32(src[key]) if src[key].dtype == torch.uint8 else src[key]
                b[key + '_len'] = src[key + '_len']
        if 'aus' in src:
            b['aus'] = src['aus']
        if 'labels' in src:
            T = meta['T_dec']
            labels = src['labels']
            go = torch.full((labels.shape[0], 1), self._decoder._GO_ID, dtype=torch.int32, device='cuda')
            b['labels'] = labels
            b['labels_len'] = src['labels_len']
            b['dec_in_ids'] = torch.cat([go, labels], dim=1)[:, :T].t().contiguous()  # decoder_unimodal.py:61-68
            b['T_dec'] = T
        self._batch = b
        return b

    # ---- forward ---------------------------------------------------------------------
    def _chains_fit(self, b):
        """Two persistent recurrent kernels side by side: 2 * 4 * ceil(B / 8) <= 148 SMs (layers.BuildContext)."""
        if not ops.tensor_cores_enabled() or self.serial_chains:
            return False
        B = int(b['labels'].shape[0]) if 'labels' in b else int(next(b[k] for k in ('audio', 'video') if k in b).shape[0])
        return 8 * ((B + 7) // 8) <= 148

    def _fork(self):
        """Side stream forked from the current one (also under graph capture): the video encoder and the
        audio layers below the cross-modal attention are independent, and a persistent LSTM kernel with
        32-utterance slices occupies only 64 of the 148 SMs, so two of them run side by side."""
        if self._side_stream is None:
            self._side_stream = torch.cuda.Stream()
        self._side_stream.wait_stream(torch.cuda.current_stream())
        return self._side_stream

    def _join(self):
        torch.cuda.current_stream().wait_stream(self._side_stream)

    def _video_features(self, b):
        """Lip crops -> CNN features [B,T,cnn_dense_units] (avsr.py:686-696), or the features as they are."""
        x = b['video']
        if self._cnn is None:
            return x
        B, T = x.shape[0], x.shape[1]
        feats = self._cnn.forward(x.reshape(B * T, *x.shape[2:]), train=self._mode == 'train')
        return feats.view(B, T, self._cnn.out_dim)

    def _video_backward(self, d, dstate):
        """Backward of the video branch: encoder, then the CNN front-end if there is one."""
        if self._cnn is None:
            self._video_encoder.backward(d, dstate)
            return
        dfeat = self._video_encoder.backward(d, dstate, need_dx=True)  # frame-major [T,B,F]
        T, B, F = dfeat.shape
        self._cnn.backward(ops.transpose01(dfeat).view(B * T, F))

    def _encode(self, b):
        enc = {}
        self._ctx.parallel_chains = self._chains_fit(b)
        both = self._video_encoder is not None and self._audio_encoder is not None
        overlap = both and (self.overlap_streams or self._ctx.parallel_chains)
        if self._video_encoder is not None:
            if overlap:
                with torch.cuda.stream(self._fork()):
                    enc['video'] = self._video_encoder.forward(self._video_features(b), b['video_len'], batch_major=True)
            else:
                enc['video'] = self._video_encoder.forward(self._video_features(b), b['video_len'], batch_major=True)
        if self._audio_encoder is not None:
            if isinstance(self._audio_encoder, AttentiveEncoder):
                self._audio_encoder.forward_lower(b['audio'], b['audio_len'], batch_major=True)
                if overlap:
                    self._join()
                enc['audio'] = self._audio_encoder.forward_top(enc['video'].outputs, b['video_len'],
                                                               enc['video'].outputs_operand)
            else:
                enc['audio'] = self._audio_encoder.forward(b['audio'], b['audio_len'], batch_major=True)
                if overlap:
                    self._join()
        return enc

    def _decoder_inputs(self, b, enc):
        if self._hparams.architecture == 'bimodal' and len(enc) == 2:
            mems = [(enc['video'].outputs, b['video_len'], enc['video'].outputs_operand),
                    (enc['audio'].outputs, b['audio_len'], enc['audio'].outputs_operand)]
            states = [enc['video'].final_state, enc['audio'].final_state]
        else:
            key = 'audio' if 'audio' in enc else 'video'
            mems = [(enc[key].outputs, b[key + '_len'], enc[key].outputs_operand)]
            states = [enc[key].final_state]
        return mems, states

    def encode(self, data_sequences=None):
        """Parity probe: encoder outputs / final states (frame-major device tensors)."""
        if data_sequences is not None:
            self.feed(data_sequences)
        return self._encode(self._prep())

    def forward_backward(self):
        """Forward + backward on the prepared batch.  Leaves the local gradient sums in store.grad and
        the cross-entropy SUM in _loss_dev[0] (normalised by the device scalar inv_denom)."""
        b = self._batch if self._batch is not None else self._prep()
        self._ctx.batch_scale = self._meta.get('batch_scale', float(self._ctx.world_size))
        self._ctx.parallel_chains = self._chains_fit(b)
        self.store.grad.zero_()
        self._loss_dev.zero_()
        if self._ctx.world_size > 1:
            # exact large-batch loss denominator (seq2seq.sequence_loss, seq2seq.py:165-171): the token count of the
            # GLOBAL batch, summed over ranks on the device inside the step (no host round trip, graph-capturable)
            tok = b['labels_len'].sum(dtype=torch.float32).reshape(1)
            if self._hparams.label_smoothing > 0.0 and self._hparams.loss_fun is None:  # smoothed loss: unmasked mean
                tok = torch.full((1,), float(b['labels_len'].shape[0] * b['T_dec']), dtype=torch.float32, device=tok.device)
            self._ctx.allreduce(tok)
            torch.reciprocal(tok + 1e-12, out=self._scal_dev[0:1])
        self._loss_dev[5:6].copy_(self._scal_dev[0:1])
        enc = self._encode(b)
        mems, states = self._decoder_inputs(b, enc)
        self._decoder.forward_train(mems, states, b['dec_in_ids'], b['labels'], b['labels_len'], b['T_dec'],
                                    self._scal_dev[0:1], self._loss_dev[0:1])
        dmem, dstates = self._decoder.backward_train()
        dvid_au = None
        if self._au_head:  # batch_loss += au_loss_weight * au_loss (seq2seq.py:188-190)
            self._video_encoder.au_loss_forward(b['aus'], self._scal_dev[2:3], self._loss_dev[3:4])
            dvid_au = self._video_encoder.au_loss_backward(None)
        both = self._video_encoder is not None and self._audio_encoder is not None
        overlap = both and (self.overlap_streams or self._ctx.parallel_chains)
        if self._hparams.architecture == 'bimodal' and both:
            if overlap:
                with torch.cuda.stream(self._fork()):
                    self._video_backward(self._plus(dmem[0], dvid_au), dstates[0])
                self._audio_encoder.backward(dmem[1], dstates[1])
                self._join()
            else:
                self._audio_encoder.backward(dmem[1], dstates[1])
                self._video_backward(self._plus(dmem[0], dvid_au), dstates[0])
        elif self._audio_encoder is not None:
            if isinstance(self._audio_encoder, AttentiveEncoder):
                d_lower, dvid = self._audio_encoder.backward_top(dmem[0], dstates[0])
                if overlap:
                    with torch.cuda.stream(self._fork()):
                        self._video_backward(self._plus(dvid, dvid_au), None)
                    self._audio_encoder.backward_lower(d_lower)
                    self._join()
                else:
                    self._audio_encoder.backward_lower(d_lower)
                    self._video_backward(self._plus(dvid, dvid_au), None)
            else:
                self._audio_encoder.backward(dmem[0], dstates[0])
        else:
            self._video_backward(self._plus(dmem[0], dvid_au), dstates[0])

    @staticmethod
    def _plus(d, extra):
        if extra is None:
            return d
        if d is None:
            return extra
        ops.axpy(1.0, extra, d)
        return d

    def _lr_now(self):
        hp = self._hparams
        lr = hp.learning_rate
        if hp.lr_decay is not None:  # seq2seq.py:263-273
            if hp.lr_decay[0] == 'cosine_restarts':
                lr = cosine_decay_restarts(lr, self._global_step, hp.lr_decay[1])
            elif not getattr(self, '_lr_policy_warned', False):
                print('learning rate policy not implemented, falling back to constant learning rate')
                self._lr_policy_warned = True
        steps = hp.kwargs.get('warmup_steps', 750)
        if steps:
            lr *= min(1.0, (self._global_step + 1) / float(steps))  # seq2seq.py:275-280
        return lr

    def finish_gradients(self):
        """all-reduce (DP) -> + L2 on the LSTM kernels -> squared global norm
        (seq2seq.py:175-178, 222, 246).  After this store.grad holds d(batch_loss)/d(theta)."""
        hp, ctx, st = self._hparams, self._ctx, self.store
        if ctx.world_size > 1:
            ctx.allreduce(st.grad)
            ctx.allreduce(self._loss_dev[0:1])
            if self._au_head:
                ctx.allreduce(self._loss_dev[3:4])
        if hp.recurrent_l2_regularisation is not None:
            for n in self._l2_names:
                ops.axpy(hp.recurrent_l2_regularisation, st.p(n), st.g(n))
                ops.sumsq(st.p(n), self._loss_dev[1:2])
        if self._cnn is not None:  # conv kernel_regularizer (video.py:27), summed into the loss by seq2seq.py:180-184
            self._cnn.add_l2(self._loss_dev[4:5])
            for enc in (self._video_encoder, self._audio_encoder):  # ... together with the rest of REGULARIZATION_LOSSES
                if enc is not None and getattr(enc, '_dense', None) is not None:
                    enc._dense.add_l2(self._loss_dev[4:5], 1e-3)
        ops.sumsq(st.grad, self._loss_dev[2:3])

    def apply_gradients(self):
        """clip_by_global_norm + TF-Adam (seq2seq.py:195-199, 245-257); lr_t is the device scalar
        _scal_dev[1] written by _set_step_scalars (warm-up + bias correction, seq2seq.py:275-280)."""
        hp, st = self._hparams, self.store
        clip = hp.max_gradient_norm if hp.clip_gradients is True else 0.0
        if hp.optimiser == 'Adam':
            ops.adam_clip_step(st.flat, st.grad, st.m, st.v, self._loss_dev[2:3], clip, self._scal_dev[1:2], 0.9, 0.999,
                               1e-8, params_tf32=st.flat_tc)
        else:
            ops.optim_clip_step(hp.optimiser, st.flat, st.grad, st.m, st.v, self._loss_dev[2:3], clip,
                                self._scal_dev[1:2], 0.9, 0.999, 1e-8,
                                weight_decay=hp.weight_decay if hp.optimiser == 'AdamW' else 0.0,
                                params_tf32=st.flat_tc)

    def _set_step_scalars(self):
        """Host scalars of this step -> device (outside any captured graph)."""
        ctx = self._ctx
        # Loss denominator: one rank knows it on the host; under data parallelism the device sums the token counts of
        # all ranks inside the step (forward_backward) and overwrites this local estimate, which then only sizes the
        # power-of-two operand scale of the backward kernels.
        n_tok = self._meta['n_tokens'] * ctx.world_size
        self._inv_denom = 1.0 / (n_tok + 1e-12)  # seq2seq.sequence_loss
        if self._hparams.label_smoothing > 0.0 and self._hparams.loss_fun is None:  # (loss_fun takes precedence, seq2seq.py:147-163)
            # seq2seq.py:147-155: smoothed_cross_entropy (devel.py:54-61) returns a reduced scalar - the mean over ALL
            # B x T positions, padding included - which sequence_loss multiplies by the weights and divides by their sum
            self._inv_denom = 1.0 / (self._meta['n_positions'] * ctx.world_size)
        self._ctx.grad_scale = float(2 ** int(math.floor(math.log2(max(n_tok, 1.0)))))
        lr = self._lr_now()
        self.current_lr = lr
        t = self._global_step + 1
        # pageable source: the driver stages it at call time, so the host may run ahead of the GPU safely
        self._au_scale = 0.0
        if self._au_head:
            cnt = parallel.global_token_count(self._meta['au_count'], device='cuda') if ctx.world_size > 1 \
                else self._meta['au_count']
            self._au_scale = float(self._hparams.kwargs.get('au_loss_weight', 10.0)) / max(cnt, 1.0)
        # Adam family: the bias correction is folded into the step size; Momentum takes the plain learning rate
        lr_t = lr if self._hparams.optimiser == 'Momentum' else lr * math.sqrt(1.0 - 0.999 ** t) / (1.0 - 0.9 ** t)
        self._scal_dev.copy_(torch.tensor([self._inv_denom, lr_t, self._au_scale], dtype=torch.float32))
        self._ctx.rng.copy_(torch.tensor(self.rng_words(), dtype=torch.int32))

    def rng_words(self):
        """(seed, step) of the generator for the CURRENT training step (31-bit so they fit the int32 device words).
        Under data parallelism every rank draws its own masks (its utterances are different ones): the rank is folded
        into the seed."""
        seed = (self.rng_seed + 0x3C6EF35F * parallel.rank()) & 0x7FFFFFFF
        return (seed, self._global_step & 0x7FFFFFFF)

    @property
    def random_streams(self):
        """cell name -> first generator stream (what the oracle needs to regenerate the masks)."""
        return dict(self._ctx.streams)

    def _step_body(self):
        self._prep()
        self.forward_backward()
        self.finish_gradients()
        self.apply_gradients()

    def _capture(self, key):
        """Capture one whole training step (thousands of launches) into a CUDA graph.  A throw-away eager
        step warms every kernel first; parameters, Adam slots and BN statistics are restored after it."""
        st = self.store
        snap = [t.clone() for t in (st.flat, st.flat_tc, st.m, st.v)] + [v.clone() for v in st.state.values()]
        n0 = ops.launch_count()
        self._step_body()
        torch.cuda.synchronize()
        for dst, src in zip([st.flat, st.flat_tc, st.m, st.v] + list(st.state.values()), snap):
            dst.copy_(src)
        n1 = ops.launch_count()
        g = torch.cuda.CUDAGraph()
        with torch.cuda.graph(g):
            self._step_body()
        self._graphs[key] = (g, n1 - n0)
        return self._graphs[key]

    def fetch_scalars(self):
        """Device -> host read of the step's results (what session.run returns, avsr.py:265-271)."""
        hp = self._hparams
        vals = self._loss_dev.cpu().numpy()
        self._inv_denom = float(vals[5])  # the denominator the device used (global token count under DP)
        xent = float(vals[0]) * self._inv_denom  # sum(xent*w) / (sum(w) + 1e-12)
        reg = 0.5 * (hp.recurrent_l2_regularisation or 0.0) * float(vals[1]) + 0.5 * 1e-3 * float(vals[4])
        self.au_loss = float(vals[3]) * self._au_scale / float(hp.kwargs.get('au_loss_weight', 10.0)) \
            if self._au_head else None
        self.batch_loss = xent + reg + (float(vals[3]) * self._au_scale if self._au_head else 0.0)
        self.global_norm = math.sqrt(float(vals[2]))
        return self.batch_loss, self.global_norm

    def fetch_scalars_async(self):
        """Starts the device -> host copy of THIS step's results into pinned memory and returns a handle whose
        `result()` waits for just that copy.  Lets the host launch step k+1 before it reads the loss of step k, so the
        GPU never waits for the host between steps (session.run pipelines the same way behind its fetches)."""
        if not hasattr(self, '_pinned_scalars'):
            self._pinned_scalars = [torch.empty(6, dtype=torch.float32).pin_memory() for _ in range(4)]
            self._pinned_next = 0
        buf = self._pinned_scalars[self._pinned_next % len(self._pinned_scalars)]
        self._pinned_next += 1
        buf.copy_(self._loss_dev, non_blocking=True)
        ev = torch.cuda.Event()
        ev.record()
        return _PendingScalars(self, buf, ev, self._inv_denom, self._au_scale)

    def train_step(self, data_sequences=None, fetch=True):
        if self._mode != 'train':
            raise Exception('train_step needs mode == `train`')
        if data_sequences is not None:
            self.feed(data_sequences)
        elif self._pending is not None:
            self._consume_prefetch()
        self._set_step_scalars()
        if self._global_step % self.BN_GAMMA_CHECK_EVERY == 0:
            self._guard_bn_shortcut()
        if self.use_cuda_graph:
            key = self._meta['key']
            g, n = self._graphs.get(key) or self._capture(key)
            g.replay()
            self.launches_last_step = n
        else:
            n0 = ops.launch_count()
            self._step_body()
            self.launches_last_step = ops.launch_count() - n0
        self._global_step += 1
        if fetch:
            return self.fetch_scalars()
        return None

    BN_GAMMA_CHECK_EVERY = 64
    BN_GAMMA_FLOOR = 0.05

    def _guard_bn_shortcut(self):
        """The input normalisation's dgamma is read off the layer-0 weight gradient divided by gamma (avsr_bn_input_grads)
        when nothing sits between the two.  A gamma that training has driven towards zero amplifies the tf32 rounding of
        that weight gradient into a spurious dgamma (and through the global norm into every clipped gradient), so every
        BN_GAMMA_CHECK_EVERY steps the smallest |gamma| is read back (one 4-byte D2H copy) and an encoder that fell under
        the floor switches - for good - to the explicit path (gradient wrt the normalised features, stored xhat).  The
        captured graphs are dropped so that the new path is what replays."""
        changed = False
        for enc in (self._video_encoder, self._audio_encoder):
            bn = getattr(enc, '_bn', None) if enc is not None else None
            if bn is None or getattr(enc, 'explicit_bn_backward', False) or enc._layer0_drops_input():
                continue
            if float(self.store.p(bn.gamma).abs().min().item()) < self.BN_GAMMA_FLOOR:
                enc.explicit_bn_backward = True
                changed = True
        if changed:
            self._graphs.clear()

    def train_op(self):
        return self.train_step()

    # ---- inference -------------------------------------------------------------------------
    def predict(self, data_sequences=None):
        """Runs the decoding algorithm of hparams (avsr.py:345): int32 ids [B, <=150]."""
        if data_sequences is not None:
            self.feed(data_sequences)
        b = self._prep()
        enc = self._encode(b)
        mems, states = self._decoder_inputs(b, enc)
        if self._hparams.decoding_algorithm == 'greedy':
            return self._decoder.decode_greedy(mems, states)
        return self._decoder.decode_beam(mems, states)
