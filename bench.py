#!/usr/bin/env python
"""bench.py - headline benchmark: AV-Align training throughput (utterances/s).

Workload (BASELINE.json configs[4], the configuration the metric is quoted on): AV-Align
cross-modal fusion, 3x256 uni-LSTM video and audio encoders, 1x256 attention decoder,
per-GPU batch 256, T_audio = 300 mel-80 frames, T_video = 75 lip crops of 36x36x3 (fed as
flat 3888-d `features`, the only video entry the six hot-path files define; the ResNet
front-end is SURVEY.md row f-3), 40-char targets + EOS.  One step = forward + backward +
global-norm clip + Adam on one batch.

  python bench.py --gpus N --steps K --warmup W          # our arm (one rank per GPU under torchrun)
  python bench.py --impl reference ...                   # the CPU restatement of the TF1 graph (oracle)

Prints ONE JSON line on rank 0.
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

import numpy as np  # noqa: E402

METRIC = 'AV-Align train utterances/sec'
UNIT = 'utterances/s'


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument('--gpus', type=int, default=1)
    ap.add_argument('--steps', type=int, default=10)
    ap.add_argument('--warmup', type=int, default=3)
    ap.add_argument('--impl', default='b200', choices=['b200', 'reference'])
    ap.add_argument('--batch', type=int, default=256, help='per-GPU batch (weak scaling)')
    ap.add_argument('--video-input', default='crops3888', choices=['crops3888', 'features128'])
    ap.add_argument('--attention', default='scaled_luong', choices=['bahdanau', 'scaled_luong'],
                    help='scorer of the cross-modal and decoder attention; scaled_luong is the reference default '
                         '(avsr.py:50) used by its AV-Align script (run_audiovisual.py:55)')
    ap.add_argument('--no-graph', action='store_true', help='eager launches instead of one CUDA graph per step')
    ap.add_argument('--no-tensor-cores', action='store_true')
    ap.add_argument('--cpu-sample', type=int, default=16, help='utterances per CPU-baseline step')
    ap.add_argument('--skip-cpu-baseline', action='store_true')
    ap.add_argument('--overlap', action='store_true', help='run the video and audio encoder branches on two streams')
    ap.add_argument('--skip-roofline', action='store_true')
    ap.add_argument('--skip-extras', action='store_true',
                    help='skip the two side measurements: the reference-default training graph (dropout + scheduled '
                         'sampling on) and the TFRecord-fed end-to-end loop')
    ap.add_argument('--tfrecord-utterances', type=int, default=1024)
    ap.add_argument('--cer-check-only', action='store_true', help=argparse.SUPPRESS)  # child process of the CER check
    return ap.parse_args()


def workload(args, B, seed):
    from tests.helpers import config_hparams, synthetic_batch
    hp = config_hparams(5, attention_type=((args.attention,), (args.attention,)))
    Fv = 3888 if args.video_input == 'crops3888' else 128
    batch = synthetic_batch(hp, B=B, Ta=300, Tv=75, Fa=80, Fv=Fv, L=40, ragged=False, seed=seed)
    return hp, batch


def config_dict(args, N):
    return {
        'workload': 'AV-Align (BASELINE.json configs[4]): 3x256 uni-LSTM video+audio encoders, cross-modal '
                    f'{args.attention} attention in the top audio layer, 1x256 {args.attention} attention decoder',
        'per_gpu_batch': args.batch, 'global_batch': args.batch * N, 'T_audio': 300, 'audio_features': 80,
        'T_video': 75, 'video_features': 3888 if args.video_input == 'crops3888' else 128,
        'video_input': '36x36x3 lip crops as flat features' if args.video_input == 'crops3888'
        else '128-d visual features', 'label_len': 41, 'parallelism': f'dp{N}',
        'dropout': 'off', 'scheduled_sampling': 'off (the parity switches of SURVEY.md 8d, the graph the reference '
                                                'arm / cpu_baseline runs too; the reference-default graph with '
                                                'both ON is timed beside it: key reference_default_graph)',
        'l2_flush': 'not needed: every step streams > 2 GB of activations through a 126 MB L2',
    }


# ------------------------------------------------------------------------------------
# reference arm / cpu baseline: the oracle (CPU restatement of the TF1 graph) on host cores
# ------------------------------------------------------------------------------------
def run_oracle(args, steps, warmup, sample):
    from avsr_tf1_b200.seq2seq import Seq2SeqModel
    from oracle import avsr_oracle as O
    from tests.helpers import oracle_hparams, to_data_sequences
    hp, batch = workload(args, sample, seed=0)
    model = Seq2SeqModel(to_data_sequences(batch), 'train', hp, seed=2001, device='cpu')
    P = model.store.to_numpy('p')
    names = model.store.names()
    om = O.OracleModel(oracle_hparams(hp), P)
    m = {k: np.zeros_like(P[k]) for k in names}
    v = {k: np.zeros_like(P[k]) for k in names}
    times = []
    for s in range(warmup + steps):
        t0 = time.perf_counter()
        loss, G, _ = om.loss_and_grads(batch)
        Pt = {k: P[k] for k in names}
        O.clip_and_adam(Pt, G, m, v, s, hp.learning_rate, clip=hp.max_gradient_norm)
        P.update(Pt)
        times.append(time.perf_counter() - t0)
    t = float(np.mean(times[warmup:]))
    try:
        import threadpoolctl
        threads = max([p['num_threads'] for p in threadpoolctl.threadpool_info()] + [1])
    except Exception:
        threads = os.cpu_count() or 1
    return dict(value=sample / t, sec_per_step=t, cores=int(threads), sample=sample, loss=float(loss))


def reference_arm(args):
    rank = int(os.environ.get('RANK', '0'))
    if rank != 0:
        return
    steps = max(1, min(args.steps, 12))
    warmup = max(1, min(args.warmup, 2))
    r = run_oracle(args, steps, warmup, args.cpu_sample)
    line = {
        'impl': 'reference', 'metric': METRIC, 'value': r['value'], 'unit': UNIT, 'n_gpus': args.gpus,
        'steps': steps, 'warmup': warmup, 'ms_per_step': r['sec_per_step'] * 1e3, 'higher_is_better': True,
        'scaling': 'weak', 'vs_baseline': None, 'dtype': 'f32', 'data': 'synthetic',
        'config': config_dict(args, args.gpus),
        'cpu_baseline': {'value': r['value'], 'unit': UNIT, 'cores': r['cores'], 'kind': 'port',
                         'sample': f'{r["sample"]} utterances per step of the same workload (full sequence '
                                   f'lengths), {steps} steps; NumPy/OpenBLAS restatement of the TF1 graph '
                                   '(TensorFlow 1.13 cannot be installed here)'},
        'e2e': {'value': r['value'], 'unit': UNIT, 'h2d_bytes_per_step': 0, 'd2h_bytes_per_step': 0},
    }
    print(json.dumps(line), flush=True)


# ------------------------------------------------------------------------------------
# clocks
# ------------------------------------------------------------------------------------
class ClockSampler(threading.Thread):
    Q = ('index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,'
         'clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,'
         'clocks_event_reasons.sw_power_cap')

    def __init__(self, index):
        super().__init__(daemon=True)
        self.index, self.rows, self.stop_flag = index, [], False

    def run(self):
        while not self.stop_flag:
            try:
                out = subprocess.run(['nvidia-smi', f'--id={self.index}', f'--query-gpu={self.Q}',
                                      '--format=csv,noheader,nounits'], capture_output=True, text=True, timeout=5)
                for ln in out.stdout.strip().splitlines():
                    self.rows.append([x.strip() for x in ln.split(',')])
            except Exception:
                pass
            time.sleep(0.1)

    def summary(self):
        sm, mx, reasons = [], 0.0, set()
        for r in self.rows:
            try:
                sm.append(float(r[1]))
                mx = max(mx, float(r[2]))
                for name, val in zip(('hw_slowdown', 'hw_thermal_slowdown', 'sw_thermal_slowdown', 'sw_power_cap'),
                                     r[5:9]):
                    if val.lower().startswith('active'):
                        reasons.add(name)
            except Exception:
                continue
        return {'sm_mhz': float(np.median(sm)) if sm else None, 'sm_max_mhz': mx or None,
                'reasons': sorted(reasons), 'samples': len(sm)}


# ------------------------------------------------------------------------------------
# roofline of the dominant kernel class: the LSTM gate GEMMs (tensor-pipe bound)
# ------------------------------------------------------------------------------------
def gate_gemm_roofline(args, torch, ops):
    """Times the big LSTM gate products of this workload in isolation with CUDA events on the launch
    stream (distinct operand buffers per launch, > L2 in total).  FLOP = 2*M*N*K per product."""
    B = args.batch
    Fv = 3888 if args.video_input == 'crops3888' else 128
    H = 256
    shapes = []  # (ta, tb, M, N, K) forward x@Wx, dgrad dZ@Wx^T, wgrad x^T@dZ for every encoder layer
    for T, I in ((75, Fv), (75, H), (75, H), (300, 80), (300, H), (300, H)):
        M = T * B
        shapes += [(0, 0, M, 4 * H, I), (0, 1, M, I, 4 * H), (1, 0, I, 4 * H, M)]
    flops, ms = 0.0, 0.0
    per = []
    for ta, tb, M, N, K in shapes:
        a = torch.randn((K, M) if ta else (M, K), device='cuda')
        b = torch.randn((N, K) if tb else (K, N), device='cuda')
        c = torch.empty(M, N, device='cuda')
        for _ in range(2):
            ops.gemm(a, b, c, ta=bool(ta), tb=bool(tb))
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        reps = 5
        torch.cuda.synchronize()
        e0.record()
        for _ in range(reps):
            ops.gemm(a, b, c, ta=bool(ta), tb=bool(tb))
        e1.record()
        torch.cuda.synchronize()
        t = e0.elapsed_time(e1) / reps
        f = 2.0 * M * N * K
        flops += f
        ms += t
        per.append({'shape': [int(ta), int(tb), M, N, K], 'ms': round(t, 4), 'tflops': round(f / t / 1e9, 2)})
        del a, b, c
    # TF32 cuBLAS peak, measured the way MEASURED_PEAKS.json measures bf16 (denominator only, not on the path)
    n = 8192
    x, y = torch.randn(n, n, device='cuda'), torch.randn(n, n, device='cuda')
    old = torch.backends.cuda.matmul.allow_tf32
    torch.backends.cuda.matmul.allow_tf32 = True
    best = 1e9
    for _ in range(6):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        torch.matmul(x, y)
        e1.record()
        torch.cuda.synchronize()
        best = min(best, e0.elapsed_time(e1))
    torch.backends.cuda.matmul.allow_tf32 = old
    tf32_peak = 2.0 * n ** 3 / best / 1e9
    peaks = {}
    try:
        peaks = json.load(open(os.path.join(ROOT, 'MEASURED_PEAKS.json')))
    except Exception:
        pass
    achieved = flops / ms / 1e9
    return {'bound': 'tensor', 'achieved': round(achieved, 2), 'peak': round(tf32_peak, 1), 'unit': 'TFLOP/s',
            'frac': round(achieved / tf32_peak, 4), 'traffic': None,
            'kernel': 'LSTM gate GEMMs (x@Wx, dZ@Wx^T, x^T@dZ of the six encoder layers), timed in isolation',
            'peak_source': 'TF32 torch.matmul 8192^3 measured in this run (operands are fp32/TF32, BASELINE.md '
                           'section 2); bf16 peak of MEASURED_PEAKS.json = %s' % peaks.get('bf16_tflops'),
            'gate_gemm_ms_per_step': round(ms, 3), 'per_shape': per}


def persistent_kernel_rooflines(args, torch, ops, model, gate):
    """The kernels that dominate the step, timed LIVE with CUDA events on their launching stream (in-library timers,
    avsr_kernel_timing) over three eager replays of the same training step that was timed above as a CUDA graph.

    Attention class (HBM): algorithmic bytes per (utterance, query step) = 4 Tm (A + Dm) + 4 (Tm + Dm + A)
    (fp32 keys + values streamed once per step); the backward kernel sweeps both again.  Tensor class (8d-1): the
    recurrent gate products 2 K 4H per (utterance, step), K = H for the plain layers, H + Dm for the attention layers,
    forward and backward alike."""
    B, H, A, Dm = args.batch, 256, 256, 256
    peaks = {}
    try:
        peaks = json.load(open(os.path.join(ROOT, 'MEASURED_PEAKS.json')))
    except Exception:
        pass
    hbm_peak = float(peaks.get('hbm_gbs', 6650.0))
    f16_peak = float(peaks.get('bf16_tflops_sustained', peaks.get('bf16_tflops', 1590.0)))
    peak_src = 'MEASURED_PEAKS.json' if peaks else 'fallback of B200_PROFILING.md'
    was_graph = model.use_cuda_graph
    model.use_cuda_graph = False
    reps = 3
    model.train_step(fetch=False)
    torch.cuda.synchronize()
    ops.kernel_timing(True)
    for _ in range(reps):
        model.train_step(fetch=False)
    torch.cuda.synchronize()
    kt = ops.kernel_times()
    ops.kernel_timing(False)
    model.use_cuda_graph = was_graph
    ms = {k: v[0] / reps for k, v in kt.items()}
    n = {k: v[1] // reps for k, v in kt.items()}

    def att_bytes(Tm):
        return 4.0 * Tm * (A + Dm) + 4.0 * (Tm + Dm + A)
    layers = ((300, 75), (41, 300))  # (query steps, memory rows): cross-modal audio layer, decoder
    bytes_dir = sum(B * T * att_bytes(Tm) for T, Tm in layers)          # per direction (fwd or bwd)
    once = sum(B * 4.0 * Tm * (A + Dm) for _, Tm in layers)              # keys + values read once per utterance
    t_att = ms['attn_lstm_fwd'] + ms['attn_lstm_bwd']
    ach = 2.0 * bytes_dir / (t_att * 1e-3) / 1e9 if t_att > 0 else 0.0
    flop_attn = sum(B * T * 2.0 * (H + Dm) * 4 * H for T, _ in layers)   # per direction
    flop_lstm = B * (2 * 300 + 3 * 75) * 2.0 * H * 4 * H                 # audio layers 0-1 + three video layers
    t_rec = t_att + ms['lstm_fwd'] + ms['lstm_bwd']
    rec_tflops = 2.0 * (flop_attn + flop_lstm) / (t_rec * 1e-3) / 1e12 if t_rec > 0 else 0.0
    roof = {
        'bound': 'hbm', 'achieved': round(ach, 1), 'peak': hbm_peak, 'unit': 'GB/s',
        'frac': round(ach / hbm_peak, 4),
        # dram__bytes_read.sum + dram__bytes_write.sum of the two launches of the cross-modal layer (forward 0.938 GB,
        # backward 0.983 GB) in profiles/r01_ncu_full_persist4.csv: the activations written for / read by the backward
        # pass.  The fp16 keys / values (19.7 MB per batch) stay in L2.
        'traffic': 1.921e9,
        'kernel': 'ap4::attn_lstm_persist4_{fwd,bwd}_kernel (cross-modal audio layer T=300/Tm=75 + decoder T=41/Tm=300)',
        'ms_per_step': {'attn_lstm_fwd': round(ms['attn_lstm_fwd'], 4), 'attn_lstm_bwd': round(ms['attn_lstm_bwd'], 4)},
        'launches_per_step': n['attn_lstm_fwd'] + n['attn_lstm_bwd'],
        'algorithmic_bytes_per_step': int(2 * bytes_dir),
        'once_per_utterance_bytes_per_step': int(2 * once),
        'peak_source': f'hbm_gbs of {peak_src}',
        'note': 'algorithmic bytes = keys + values streamed once per query step as fp32 (SURVEY.md 8d); the kernels read '
                'fp16 copies (half of it) and the memories stay resident in the 126 MB L2, so the figure measures L2-fed '
                'sweeps against the HBM peak: the layer is bound by the per-step latency chain exchange -> product -> '
                'gate math -> sweep, not by DRAM (ncu: dram throughput 5 %, issue slots 37 %, tensor pipe 4 %)',
        'timing': f'CUDA events around each launch on its stream, {reps} eager steps after the graph-timed loop',
    }
    tensor = {
        'bound': 'tensor', 'unit': 'TFLOP/s',
        'recurrent_products': {
            'achieved': round(rec_tflops, 2), 'peak': f16_peak, 'frac': round(rec_tflops / f16_peak, 5),
            'kernels': 'lp4::lstm_persist4_{fwd,bwd}_kernel + ap4::attn_lstm_persist4_{fwd,bwd}_kernel (fp16 operands, '
                       'fp32 accumulation in TMEM)',
            'ms_per_step': {k: round(ms[k], 4) for k in ('lstm_fwd', 'lstm_bwd', 'attn_lstm_fwd', 'attn_lstm_bwd')},
            'flop_per_step': 2.0 * (flop_attn + flop_lstm),
            'peak_source': f'bf16 sustained of {peak_src}',
            'note': 'one 128 x 16 x 16 product chain per time step and CTA between two cluster exchanges: latency bound by '
                    'construction (sequential recurrence), the tensor pipe is ~4-6 % busy (ncu)'},
        'gemm_kernel_ms_per_step': round(ms['gemm'], 4), 'gemm_launches_per_step': n['gemm'],
        'gate_gemms': gate,
    }
    return roof, tensor


def default_graph_throughput(args, torch, ds_host, steps):
    """The reference's DEFAULT training graph (avsr.py:49-56: DropoutWrapper keep 0.9 on input / state / output of
    every cell, scheduled sampling 0.1) on the same workload.  Plain LSTM layers keep their persistent kernels (masks
    regenerated in-kernel); the two attention layers run step-wise (a mask sits between the attention layer and the
    recurrent matrix, so the folded recurrence of the persistent attention kernel does not apply)."""
    from avsr_tf1_b200.seq2seq import Seq2SeqModel
    from tests.helpers import config_hparams
    hp = config_hparams(5, attention_type=((args.attention,), (args.attention,)), use_dropout=True,
                        sampling_probability_outputs=0.1)
    model = Seq2SeqModel(ds_host, 'train', hp, seed=2001)
    model.use_cuda_graph = not args.no_graph
    model.feed(ds_host)
    for _ in range(3):
        model.train_step(fetch=False)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(steps):
        model.train_step(fetch=False)
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / steps
    loss, gnorm = model.fetch_scalars()
    return {'value': round(args.batch / (ms / 1e3), 2), 'unit': UNIT, 'ms_per_step': round(ms, 4), 'steps': steps,
            'gpu_launches_per_step': int(model.launches_last_step), 'loss': round(float(loss), 6),
            'config': 'same workload with use_dropout=True (keep 0.9/0.9/0.9 on every cell) and '
                      'sampling_probability_outputs=0.1: the reference defaults; masks and draws from the '
                      'counter-based generator (include/avsr_b200.h avsr_dropout / avsr_sched_sample)'}


def tfrecord_e2e(args, torch, model, n_utt):
    """the same training step fed from synthetic TFRecords in the reference's schema (8d) through
    the native reader (include/avsr_io.h) - shuffle, bucket, padded batch into pinned memory on a prefetch thread -
    instead of one resident pinned batch."""
    import shutil
    import tempfile

    from avsr_tf1_b200 import io_utils
    from avsr_tf1_b200.synthetic import write_synthetic_records
    d = tempfile.mkdtemp(prefix='avsr_records_')
    try:
        t0 = time.perf_counter()
        paths = write_synthetic_records(d, n=n_utt, Ta=300, Tv=75, Fa=80, hw=36, channels=3, L=40)
        t_write = time.perf_counter() - t0
        cores = max(1, min(16, (os.cpu_count() or 4) - 1))
        it = io_utils.make_iterator_from_two_records(
            paths['video'], paths['audio'], paths['labels'], batch_size=args.batch,
            unit_dict=model._hparams.unit_dict, shuffle=True, bucket_width=45, num_cores=cores, prefetch=3)
        bytes_on_disk = sum(os.path.getsize(p) for p in paths.values())

        def epoch():
            """One pass over the records; the H2D copy of batch k+1 (copy stream) overlaps the step of batch k."""
            n = 0
            batches = iter(it)
            b = next(batches, None)
            if b is not None:
                model.prefetch(b.data_sequences())
            pending = None
            while b is not None:
                size = b.labels.shape[0]
                model.train_step(fetch=False)  # consumes the staged batch
                handle = model.fetch_scalars_async()  # D2H loss of this step
                b = next(batches, None)
                if b is not None:
                    model.prefetch(b.data_sequences())
                if pending is not None:
                    pending.result()  # read step k-1 while step k runs (staging buffers are reused 6 batches later)
                pending = handle
                n += size
            if pending is not None:
                pending.result()
            return n
        epoch()  # page cache, pinned staging ring and static device buffers warm
        epoch()
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        n = epoch() + epoch() + epoch()
        torch.cuda.synchronize()
        dt = time.perf_counter() - t0
        # reader alone (no GPU work): what the host side sustains
        t0 = time.perf_counter()
        m = sum(b.labels.shape[0] for b in it)
        dt_read = time.perf_counter() - t0
        return {'value': round(n / dt, 2), 'unit': UNIT, 'utterances': n, 'reader_only_utterances_per_s': round(m / dt_read, 1),
                'reader_threads': cores, 'record_bytes': int(bytes_on_disk), 'write_s': round(t_write, 2),
                'timing': 'wall clock over three epochs incl. record decode, H2D (overlapped with the previous step), step, '
                          'D2H loss every step'}
    finally:
        shutil.rmtree(d, ignore_errors=True)


def cer_check(args, torch, ops):
    """Second half of BASELINE.json's metric ("CER match vs TF1 ref"): greedy decoding of a small AV-Align batch by the
    product (exact-fp32 mode, so arg-max decisions are reproducible) against the CPU restatement of the TF1 graph; the
    predicted ids must be equal, hence the integer edit distances and the CER.  Part of the cpu_baseline leg (the only
    place besides the reference arm where bench.py runs the oracle)."""
    from avsr_tf1_b200 import utils
    from avsr_tf1_b200.seq2seq import Seq2SeqModel
    from oracle import avsr_oracle as O
    from tests.helpers import cast_batch, config_hparams, oracle_hparams, synthetic_batch, to_data_sequences
    old = ops.set_tensor_cores(False)
    try:
        hp = config_hparams(5, attention_type=((args.attention,), (args.attention,)), decoding_algorithm='greedy')
        hp.max_label_length = 12
        # (the configuration of tests/test_gpu_model.py::test_decoding_and_error_rates[5-greedy])
        batch = synthetic_batch(hp, B=3, Ta=30, Tv=10, L=6, ragged=True)
        ds = to_data_sequences(batch)
        train = Seq2SeqModel(ds, 'train', hp, seed=2001)
        train.store.p('Decoder/decoder/my_dense/kernel').mul_(20.0)  # sharpen the outputs: no near-ties
        train.store.sync_tf32()
        for _ in range(2):
            train.train_step(ds)
        ev = Seq2SeqModel(ds, 'evaluate', hp, share_params_with=train)
        ids = ev.predict(ds)
        P = {k: v.astype(np.float32) for k, v in train.store.to_numpy('p').items()}
        ref = O.OracleModel(oracle_hparams(hp), P).greedy_decode(cast_batch(batch, np.float32))
        ud = hp.unit_dict
        names = ['utt%d' % b for b in range(ids.shape[0])]
        truth = {n: utils.ids_to_symbols(batch['labels'][b], ud) for b, n in enumerate(names)}
        cer = utils.compute_wer({n: utils.ids_to_symbols(ids[b], ud) for b, n in enumerate(names)}, truth)[0]
        cer_ref = O.compute_wer({n: utils.ids_to_symbols(ref[b], ud) for b, n in enumerate(names)}, truth)[0]
        same = ids.shape == ref.shape and bool(np.array_equal(ids, ref))
        return {'ids_equal': same, 'cer': cer, 'cer_oracle': cer_ref, 'cer_equal': cer == cer_ref,
                'utterances': int(ids.shape[0]), 'decoded_steps': int(ids.shape[1]),
                'how': 'greedy decoding, exact-fp32 mode, AV-Align 3x256, after two training steps; oracle = CPU '
                       'restatement of the TF1 graph (TensorFlow 1.13 cannot run here)'}
    finally:
        ops.set_tensor_cores(old)


def main():
    args = parse()
    if args.impl == 'reference':
        reference_arm(args)
        return
    if args.cer_check_only:
        import torch
        from avsr_tf1_b200 import ops
        torch.cuda.set_device(0)
        print(json.dumps(cer_check(args, torch, ops)), flush=True)
        return
    import faulthandler

    import torch
    import torch.distributed as dist

    # a hang (e.g. a mismatched collective) must end with a traceback, not with the driver's timeout
    faulthandler.dump_traceback_later(int(os.environ.get('AVSR_BENCH_WATCHDOG_S', '900')), exit=True)
    world = int(os.environ.get('WORLD_SIZE', '1'))
    rank = int(os.environ.get('RANK', '0'))
    local = int(os.environ.get('LOCAL_RANK', '0'))
    torch.cuda.set_device(local)
    if world > 1:
        # NCCL announces its version on stdout at init; keep stdout for the ONE JSON line
        sys.stdout.flush()
        saved = os.dup(1)
        os.dup2(2, 1)
        try:
            dist.init_process_group('nccl', device_id=torch.device('cuda', local))
            dist.barrier()
            torch.cuda.synchronize()
        finally:
            sys.stdout.flush()
            os.dup2(saved, 1)
            os.close(saved)
    from avsr_tf1_b200 import ops
    from avsr_tf1_b200.seq2seq import Seq2SeqModel
    from tests.helpers import to_data_sequences

    ops.set_tensor_cores(not args.no_tensor_cores)
    hp, batch = workload(args, args.batch, seed=rank)
    pinned = {k: torch.from_numpy(v).pin_memory() for k, v in batch.items()}
    ds_host = to_data_sequences(pinned)
    model = Seq2SeqModel(ds_host, 'train', hp, seed=2001)
    model.use_cuda_graph = not args.no_graph
    model.overlap_streams = bool(args.overlap)

    def barrier():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def max_over_ranks(ms):
        if world == 1:
            return ms
        t = torch.tensor([ms], dtype=torch.float64, device='cuda')
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    # ---- device-resident timing: inputs already in HBM ------------------------------------
    model.feed(ds_host)
    for _ in range(max(3, args.warmup)):
        model.train_step(fetch=False)
    barrier()
    sampler = ClockSampler(local)
    sampler.start()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(args.steps):
        model.train_step(fetch=False)
    e1.record()
    barrier()
    ms_total = max_over_ranks(e0.elapsed_time(e1))
    launches = model.launches_last_step
    loss, gnorm = model.fetch_scalars()

    # ---- end to end: pinned host batch -> H2D -> step -> D2H loss, every step ---------------
    # The H2D copy of step k+1's batch is issued (copy stream) right after step k is launched, so it overlaps
    # the compute of step k - the input-pipeline prefetch the reference gets from tf.data (io_utils.py:145).
    # Every step still copies its full batch from pinned host memory and reads its loss back (the read of step k is
    # waited for after step k+1 has been launched, so the host round trip is off the GPU's critical path).
    for _ in range(2):
        model.train_step(ds_host, fetch=True)
    barrier()
    e2, e3 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    t_wall = time.perf_counter()
    e2.record()
    model.prefetch(ds_host)
    pending = None
    for i in range(args.steps):
        model.train_step(fetch=False)          # consumes the prefetched batch, launches the step
        handle = model.fetch_scalars_async()   # D2H of THIS step's loss / grad norm into pinned memory
        if i + 1 < args.steps:
            model.prefetch(ds_host)            # next batch's H2D overlaps this step
        if pending is not None:
            loss, gnorm = pending.result()     # the host reads step i-1 while step i runs: the GPU never idles
        pending = handle
    loss, gnorm = pending.result()
    e3.record()
    barrier()
    e2e_wall = (time.perf_counter() - t_wall) * 1e3
    e2e_ms = max_over_ranks(max(e2.elapsed_time(e3), 0.0))
    sampler.stop_flag = True
    sampler.join(timeout=2)
    h2d = model.h2d_bytes
    d2h = int(model._loss_dev.numel() * 4)

    def finish():
        # captured graphs hold NCCL work; tearing the process group down under them can hang, so leave hard
        sys.stdout.flush()
        sys.stderr.flush()
        if world > 1:
            dist.barrier()
            torch.cuda.synchronize()
            os._exit(0)

    if rank != 0:
        finish()
        return

    ms_step = ms_total / args.steps
    value = args.batch * world / (ms_step / 1e3)
    e2e_step = e2e_ms / args.steps
    line = {
        'metric': METRIC, 'value': round(value, 2), 'unit': UNIT, 'n_gpus': world, 'steps': args.steps,
        'warmup': max(3, args.warmup), 'ms_per_step': round(ms_step, 4), 'higher_is_better': True,
        'scaling': 'weak', 'vs_baseline': None,
        'dtype': 'f32' if args.no_tensor_cores else 'f32 storage/accumulate, tf32 tensor-core products',
        'data': 'synthetic', 'config': config_dict(args, world),
        'clocks': sampler.summary(),
        'e2e': {'value': round(args.batch * world / (e2e_step / 1e3), 2), 'unit': UNIT, 'ms_per_step': round(e2e_step, 4),
                'wall_ms_per_step': round(e2e_wall / args.steps, 4), 'h2d_bytes_per_step': int(h2d),
                'd2h_bytes_per_step': d2h},
        'gpu_launches': int(launches * args.steps),
        'gpu_launches_per_step': int(launches),
        'cuda_graph': bool(model.use_cuda_graph),
        'loss': round(float(loss), 6), 'global_norm': round(float(gnorm), 6), 'n_params': int(model.n_params),
    }
    try:
        if args.skip_roofline or world > 1:
            line['roofline'] = None
        else:
            gate = gate_gemm_roofline(args, torch, ops)
            line['roofline'], line['roofline_tensor'] = persistent_kernel_rooflines(args, torch, ops, model, gate)
    except Exception as ex:  # keep the headline number even if the side measurement fails
        line['roofline'] = {'error': repr(ex)}
    if world == 1 and not args.skip_extras:
        try:
            line['reference_default_graph'] = default_graph_throughput(args, torch, ds_host, max(2, min(args.steps, 5)))
        except Exception as ex:
            line['reference_default_graph'] = {'error': repr(ex)}
        try:
            line['e2e_tfrecord'] = tfrecord_e2e(args, torch, model, args.tfrecord_utterances)
        except Exception as ex:
            line['e2e_tfrecord'] = {'error': repr(ex)}
    if world == 1 and not args.skip_cpu_baseline:
        r = run_oracle(args, steps=3, warmup=1, sample=args.cpu_sample)
        line['cpu_baseline'] = {
            'value': round(r['value'], 3), 'unit': UNIT, 'cores': r['cores'], 'kind': 'port',
            'sample': f'{r["sample"]} utterances per step of the same workload at full sequence lengths, 3 timed '
                      'steps (about 10 s of host work); NumPy/OpenBLAS restatement of the TF1 graph (oracle/)'}
        try:  # in a child process: nothing it does can cost the parent its JSON line
            child = subprocess.run([sys.executable, os.path.abspath(__file__), '--cer-check-only', '--attention',
                                    args.attention], capture_output=True, text=True, timeout=300)
            out_lines = [ln for ln in child.stdout.strip().splitlines() if ln.startswith('{')]
            line['cpu_baseline']['cer_check'] = json.loads(out_lines[-1]) if out_lines else {
                'error': 'exit %d: %s' % (child.returncode, child.stderr.strip()[-300:])}
        except Exception as ex:
            line['cpu_baseline']['cer_check'] = {'error': repr(ex)}
    print(json.dumps(line), flush=True)
    finish()


if __name__ == '__main__':
    main()
