#!/usr/bin/env python
"""bench.py - headline benchmark: AV-Align training throughput (utterances/s) on the reference's DEFAULT graph.

Workload (BASELINE.json configs[4], the configuration the metric is quoted on): AV-Align cross-modal fusion, 3x256
uni-LSTM video and audio encoders, 1x256 attention decoder, per-GPU batch 256, T_audio = 300 mel-80 frames, T_video =
75 lip crops of 36x36x3 (fed as flat 3888-d `features`, the only video entry the six hot-path files define; the ResNet
front-end is SURVEY.md row f-3), 40-char targets + EOS, with the reference's default randomness ON (avsr.py:51-56:
DropoutWrapper keep 0.9 / 0.9 / 0.9 on every cell, scheduled sampling 0.1) - what run_audiovisual.py trains.  One step =
forward + backward + global-norm clip + Adam on one batch.

  python bench.py --gpus N --steps K --warmup W          # our arm (one rank per GPU under torchrun)
  python bench.py --impl reference ...                   # the CPU restatement of the TF1 graph (oracle), same graph
  python bench.py --config 2 ...                         # another BASELINE.json configuration as the workload
  python bench.py --graph parity ...                     # randomness off (the switches of the parity tests)
  python bench.py --scaling strong --gpus N              # global batch fixed at the configuration's (256 / N per GPU)

The default run also measures, as side keys of the same JSON line (N = 1): `parity_graph`, `configs` (BASELINE configs
1-4 with their kernel-class times), `e2e_tfrecord`, `cpu_baseline`; at N > 1: `strong_scaling`.  Prints ONE JSON line on
rank 0.
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

UNIT = 'utterances/s'
# BASELINE.json configs: (per-GPU batch at N = 1, label, uses video, uses audio)
CONFIGS = {
    1: dict(batch=2, what='audio-only LAS, 1x128 uni-LSTM encoder, 1x128 decoder (run_audio.py reduced; the CPU-runnable case)'),
    2: dict(batch=64, what='audio-only LAS, 3x256 BiLSTM encoder + Bahdanau attention decoder, T_a=300'),
    3: dict(batch=64, what='video-only lipreading, 3x256 LSTM over 36x36x3 lip crops (flat 3888-d features), T_v=75'),
    4: dict(batch=128, what='WLAS dual-attention AV decoder, 3x256 encoder per modality'),
    5: dict(batch=256, what='AV-Align cross-modal fusion, 3x256 encoders, T_a=300 / T_v=75'),
}
METRICS = {1: 'LAS (1x128) train utterances/sec', 2: 'LAS BiLSTM train utterances/sec', 3: 'lipreading train utterances/sec',
           4: 'WLAS train utterances/sec', 5: 'AV-Align train utterances/sec'}


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument('--gpus', type=int, default=1)
    ap.add_argument('--steps', type=int, default=10)
    ap.add_argument('--warmup', type=int, default=3)
    ap.add_argument('--impl', default='b200', choices=['b200', 'reference'])
    ap.add_argument('--config', type=int, default=5, choices=sorted(CONFIGS), help='BASELINE.json configuration (1-5)')
    ap.add_argument('--graph', default='default', choices=['default', 'parity'],
                    help='default: the reference defaults (dropout 0.9 on every cell, scheduled sampling 0.1); '
                         'parity: randomness off (SURVEY.md 8d parity switches)')
    ap.add_argument('--scaling', default='weak', choices=['weak', 'strong'],
                    help='weak: per-GPU batch fixed; strong: global batch fixed (per-GPU batch = batch / N)')
    ap.add_argument('--batch', type=int, default=0, help='per-GPU batch at N = 1 (default: the configuration\'s)')
    ap.add_argument('--video-input', default='crops3888', choices=['crops3888', 'features128'])
    ap.add_argument('--attention', default=None, choices=['bahdanau', 'scaled_luong'],
                    help='scorer override; default: scaled_luong (avsr.py:50) except config 2 (Bahdanau, BASELINE.json)')
    ap.add_argument('--no-graph', action='store_true', help='eager launches instead of one CUDA graph per step')
    ap.add_argument('--no-tensor-cores', action='store_true')
    ap.add_argument('--cpu-sample', type=int, default=16, help='utterances per CPU-baseline step')
    ap.add_argument('--skip-cpu-baseline', action='store_true')
    ap.add_argument('--overlap', action='store_true', help='run the video and audio encoder branches on two streams')
    ap.add_argument('--skip-roofline', action='store_true')
    ap.add_argument('--skip-extras', action='store_true',
                    help='skip the side measurements (parity graph, configs 1-4, TFRecord-fed loop, strong scaling)')
    ap.add_argument('--tfrecord-utterances', type=int, default=2048)
    ap.add_argument('--cer-check-only', action='store_true', help=argparse.SUPPRESS)  # child process of the CER check
    return ap.parse_args()


def randomness(graph):
    return dict(use_dropout=True, sampling_probability_outputs=0.1) if graph == 'default' else {}


def workload(args, cfg, B, seed, graph):
    """hparams + one synthetic batch (numpy) of BASELINE config `cfg`.  Lip crops are generated as PIXELS: the float
    features are (v - 128) / 128 of uint8 values (dataset_writer.py:537), `video_u8` holds the bytes."""
    from tests.helpers import config_hparams, synthetic_batch
    over = dict(randomness(graph))
    att = args.attention or ('bahdanau' if cfg == 2 else 'scaled_luong')
    over['attention_type'] = ((att,), (att,))
    hp = config_hparams(cfg, **over)
    Fv = 3888 if args.video_input == 'crops3888' else 128
    batch = synthetic_batch(hp, B=B, Ta=300, Tv=75, Fa=80, Fv=Fv, L=40, ragged=False, seed=seed)
    if 'video' in batch and Fv == 3888:
        u8 = np.clip(np.rint(batch['video'] * 128.0 + 128.0), 0, 255).astype(np.uint8)
        batch['video'] = ((u8.astype(np.float32) - 128.0) / 128.0).astype(np.float32)
        batch['video_u8'] = u8
    return hp, batch, att


def config_dict(args, cfg, B, N, att, graph, scaling):
    d = {
        'workload': f'BASELINE.json configs[{cfg - 1}]: {CONFIGS[cfg]["what"]}; {att} attention',
        'per_gpu_batch': B, 'global_batch': B * N, 'label_len': 41, 'parallelism': f'dp{N}', 'scaling': scaling,
        'dropout': 'DropoutWrapper keep 0.9/0.9/0.9 on every cell (cells.py:46-54, avsr.py:51-54)' if graph == 'default' else 'off',
        'scheduled_sampling': '0.1 (decoder_unimodal.py:304-309, avsr.py:56)' if graph == 'default' else 'off',
        'graph': 'reference default (what run_audiovisual.py trains)' if graph == 'default'
                 else 'parity switches of SURVEY.md 8d (randomness off)',
        'l2_flush': 'not needed: every step streams > 2 GB of activations through a 126 MB L2',
    }
    if cfg != 3:
        d.update(T_audio=300, audio_features=80)
    if cfg >= 3:
        d.update(T_video=75, video_features=3888 if args.video_input == 'crops3888' else 128,
                 video_input='36x36x3 lip crops as flat features; uint8 pixels over PCIe in the e2e loop, expanded to '
                             '(v-128)/128 on the device' if args.video_input == 'crops3888' else '128-d visual features')
    return d


# ------------------------------------------------------------------------------------
# reference arm / cpu baseline: the oracle (CPU restatement of the TF1 graph) on host cores
# ------------------------------------------------------------------------------------
def host_threads():
    """torchrun exports OMP_NUM_THREADS=1 to its workers; the CPU arm is meant to use every host core."""
    n = len(os.sched_getaffinity(0)) if hasattr(os, 'sched_getaffinity') else (os.cpu_count() or 1)
    try:
        import threadpoolctl
        threadpoolctl.threadpool_limits(limits=n)
        got = max([p['num_threads'] for p in threadpoolctl.threadpool_info()] + [1])
    except Exception:
        got = n
    return int(got)


def run_oracle(args, cfg, graph, steps, warmup, sample):
    from avsr_tf1_b200.seq2seq import Seq2SeqModel
    from oracle import avsr_oracle as O
    from tests.helpers import oracle_hparams, to_data_sequences
    threads = host_threads()
    hp, batch, _ = workload(args, cfg, sample, seed=0, graph=graph)
    batch.pop('video_u8', None)
    model = Seq2SeqModel(to_data_sequences(batch), 'train', hp, seed=2001, device='cpu')
    P = model.store.to_numpy('p')
    names = model.store.names()
    m = {k: np.zeros_like(P[k]) for k in names}
    v = {k: np.zeros_like(P[k]) for k in names}
    times = []
    for s in range(warmup + steps):
        t0 = time.perf_counter()
        model._global_step = s  # fresh masks / draws every step, as in training
        om = O.OracleModel(oracle_hparams(hp, model if graph == 'default' else None), P)
        loss, G, _ = om.loss_and_grads(batch)
        Pt = {k: P[k] for k in names}
        O.clip_and_adam(Pt, G, m, v, s, hp.learning_rate, clip=hp.max_gradient_norm)
        P.update(Pt)
        times.append(time.perf_counter() - t0)
    t = float(np.mean(times[warmup:]))
    return dict(value=sample / t, sec_per_step=t, cores=threads, sample=sample, loss=float(loss))


def train_flop_per_utterance(cfg, Fv):
    """2 M N K of every matrix product of one forward pass per utterance (gate products, memory layers, attention
    layers, score / context sweeps, output layer), times 3 for training."""
    H = 128 if cfg == 1 else 256
    Ta, Tv, L, E, V = 300, 75, 41, 128, 31
    f = 0.0
    def lstm(T, I, n=1):
        return n * T * 2.0 * (I + H) * 4 * H
    def attn(Tq, mems, x):  # AttentionWrapper layer: cell over [x, attention, h]; per memory: memory layer, scores + contexts, attention layer
        f = Tq * 2.0 * (x + len(mems) * H + H) * 4 * H
        for Tm, Dm in mems:
            f += Tm * 2.0 * Dm * H + Tq * (2.0 * Tm * H + 2.0 * Tm * Dm) + Tq * 2.0 * (H + Dm) * H
        return f
    if cfg == 1:
        f = lstm(Ta, 80) + attn(L, [(Ta, H)], E)
    elif cfg == 2:
        f = 2 * (lstm(Ta, 80) + lstm(Ta, H, 2)) + attn(L, [(Ta, 2 * H)], E)
    elif cfg == 3:
        f = lstm(Tv, Fv) + lstm(Tv, H, 2) + attn(L, [(Tv, H)], E)
    elif cfg == 4:
        f = lstm(Tv, Fv) + lstm(Tv, H, 2) + lstm(Ta, 80) + lstm(Ta, H, 2) + attn(L, [(Tv, H), (Ta, H)], E)
    else:
        f = lstm(Tv, Fv) + lstm(Tv, H, 2) + lstm(Ta, 80) + lstm(Ta, H) + attn(Ta, [(Tv, H)], H) + attn(L, [(Ta, H)], E)
    f += L * 2.0 * H * V
    return 3.0 * f


def reference_arm(args):
    rank = int(os.environ.get('RANK', '0'))
    if rank != 0:
        return
    cfg = args.config
    sample = min(args.cpu_sample, (args.batch or CONFIGS[cfg]['batch']))
    r = run_oracle(args, cfg, args.graph, args.steps, args.warmup, sample)
    _, _, att = workload(args, cfg, 1, 0, args.graph)
    line = {
        'impl': 'reference', 'metric': METRICS[cfg], 'value': r['value'], 'unit': UNIT, 'n_gpus': args.gpus,
        'steps': args.steps, 'warmup': args.warmup, 'ms_per_step': r['sec_per_step'] * 1e3, 'higher_is_better': True,
        'scaling': args.scaling, 'vs_baseline': None, 'dtype': 'f32', 'data': 'synthetic',
        'config': config_dict(args, cfg, args.batch or CONFIGS[cfg]['batch'], args.gpus, att, args.graph, args.scaling),
        'cpu_baseline': {'value': r['value'], 'unit': UNIT, 'cores': r['cores'], 'kind': 'port',
                         'sample': f'{r["sample"]} utterances per step of the same workload and graph (full sequence '
                                   f'lengths), {args.steps} steps; NumPy/OpenBLAS restatement of the TF1 graph '
                                   '(TensorFlow 1.13 cannot be installed here); OMP_NUM_THREADS of the launcher overridden '
                                   f'to {r["cores"]} threads'},
        'e2e': {'value': r['value'], 'unit': UNIT, 'h2d_bytes_per_step': 0, 'd2h_bytes_per_step': 0},
    }
    fl = train_flop_per_utterance(cfg, 3888 if args.video_input == 'crops3888' else 128)
    line['cpu_baseline']['gflops'] = round(r['value'] * fl / 1e9, 1)
    line['cpu_baseline']['gflop_per_utterance'] = round(fl / 1e9, 2)
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


def load_peaks():
    try:
        return json.load(open(os.path.join(ROOT, 'MEASURED_PEAKS.json')))
    except Exception:
        return {}


# ------------------------------------------------------------------------------------
# rooflines
# ------------------------------------------------------------------------------------
def gate_gemm_roofline(args, torch, ops, B):
    """Times the big LSTM gate products of the AV-Align workload in isolation with CUDA events on the launch
    stream (distinct operand buffers per launch, > L2 in total).  FLOP = 2*M*N*K per product."""
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
        per.append({'shape': [int(ta), int(tb), M, N, K], 'ms': round(t, 4), 'tflops': round(f / t / 1e9, 2),
                    'bytes': 4.0 * (M * K + K * N + M * N)})
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
    del x, y
    tf32_peak = 2.0 * n ** 3 / best / 1e9
    peaks = load_peaks()
    achieved = flops / ms / 1e9
    # every product against ITS roof: the K = 256 / 80 shapes are bound by writing the fp32 x-projection (the [T B, 4H]
    # output is 4x the input), not by the tensor pipe - time at the roof = max(flop / tensor peak, bytes / HBM peak)
    hbm = float(peaks.get('hbm_gbs', 6650.0))
    roof_ms = 0.0
    for e in per:
        ta_, tb_, M_, N_, K_ = e['shape']
        t_tensor = 2.0 * M_ * N_ * K_ / (tf32_peak * 1e9)
        t_hbm = e.pop('bytes') / (hbm * 1e6)
        e['bound'] = 'tensor' if t_tensor >= t_hbm else 'hbm'
        e['frac_of_roof'] = round(max(t_tensor, t_hbm) / e['ms'], 3)
        roof_ms += max(t_tensor, t_hbm)
    return {'bound': 'tensor', 'achieved': round(achieved, 2), 'peak': round(tf32_peak, 1), 'unit': 'TFLOP/s',
            'frac': round(achieved / tf32_peak, 4), 'traffic': None,
            'kernel': 'LSTM gate GEMMs (x@Wx, dZ@Wx^T, x^T@dZ of the six encoder layers), timed in isolation',
            'peak_source': 'TF32 torch.matmul 8192^3 measured in this run (operands are fp32/TF32, BASELINE.md '
                           'section 2); bf16 peak of MEASURED_PEAKS.json = %s' % peaks.get('bf16_tflops'),
            'frac_of_roofline': round(roof_ms / ms, 4),
            'frac_of_roofline_note': 'sum over the products of max(flop / TF32 peak, compulsory bytes / HBM peak) over the '
                                     'measured time: 12 of the 18 products are HBM-bound (fp32 gate pre-activations out)',
            'gate_gemm_ms_per_step': round(ms, 3), 'per_shape': per}


def kernel_class_times(torch, ops, model, reps=3):
    """The persistent kernels and the GEMM class, timed LIVE with CUDA events on their launching stream (in-library
    timers, avsr_kernel_timing) over `reps` eager replays of the training step that was timed as a CUDA graph."""
    was_graph = model.use_cuda_graph
    model.use_cuda_graph = False
    model.train_step(fetch=False)
    torch.cuda.synchronize()
    ops.kernel_timing(True)
    for _ in range(reps):
        model.train_step(fetch=False)
    torch.cuda.synchronize()
    kt = ops.kernel_times()
    ops.kernel_timing(False)
    model.use_cuda_graph = was_graph
    return ({k: v[0] / reps for k, v in kt.items()}, {k: v[1] // reps for k, v in kt.items()})


# dram__bytes_read.sum + dram__bytes_write.sum of one launch each of the cross-modal layer's forward and backward kernels
# (ncu --set full, profiles/): the activations written for / read by the backward pass; the fp16 keys / values stay in L2
MEASURED_DRAM_BYTES = {'parity': (1.921e9, 'profiles/r01_ncu_full_persist4.csv'),
                       'default': (2.075e9, 'profiles/r02_ncu_full_persist4d.csv (forward 0.337 + 0.756 GB, backward 0.519 + 0.463 GB)')}


def attention_roofline(args, B, graph, ms, n, gate):
    """Attention class (HBM): algorithmic bytes per (utterance, query step) = 4 Tm (A + Dm) + 4 (Tm + Dm + A)
    (fp32 keys + values streamed once per step); the backward kernel sweeps both again.  Three readings of the same
    time are given because the memories stay on chip (L2): streamed bytes (`frac`, the contract's definition), keys +
    values once per utterance, and the DRAM bytes ncu measured.  Tensor class (8d-1): the recurrent gate products
    2 K 4H per (utterance, step), K = H for the plain layers, H + Dm (+ the attention layer under dropout)."""
    H, A, Dm = 256, 256, 256
    peaks = load_peaks()
    hbm_peak = float(peaks.get('hbm_gbs', 6650.0))
    f16_peak = float(peaks.get('bf16_tflops_sustained', peaks.get('bf16_tflops', 1590.0)))
    peak_src = 'MEASURED_PEAKS.json' if peaks else 'fallback of B200_PROFILING.md'

    def att_bytes(Tm):
        return 4.0 * Tm * (A + Dm) + 4.0 * (Tm + Dm + A)
    layers = ((300, 75), (41, 300))  # (query steps, memory rows): cross-modal audio layer, decoder
    bytes_dir = sum(B * T * att_bytes(Tm) for T, Tm in layers)          # per direction (fwd or bwd)
    once = sum(B * 4.0 * Tm * (A + Dm) for _, Tm in layers)              # keys + values read once per utterance
    t_att = ms['attn_lstm_fwd'] + ms['attn_lstm_bwd']
    ach = 2.0 * bytes_dir / (t_att * 1e-3) / 1e9 if t_att > 0 else 0.0
    ach_once = 2.0 * once / (t_att * 1e-3) / 1e9 if t_att > 0 else 0.0
    dram, dram_src = MEASURED_DRAM_BYTES[graph]
    k_att = (H + Dm + (H + Dm) * A / (4.0 * H)) if graph == 'default' else (H + Dm)  # + a_t = [ho|ctx] Wa under dropout
    flop_attn = sum(B * T * 2.0 * k_att * 4 * H for T, _ in layers)      # per direction
    flop_lstm = B * (2 * 300 + 3 * 75) * 2.0 * H * 4 * H                 # audio layers 0-1 + three video layers
    t_rec = t_att + ms['lstm_fwd'] + ms['lstm_bwd']
    rec_tflops = 2.0 * (flop_attn + flop_lstm) / (t_rec * 1e-3) / 1e12 if t_rec > 0 else 0.0
    steps_total = sum(T for T, _ in layers)
    kname = 'ap4::attn_lstm_persist4d_{fwd,bwd}_kernel (two products per step: DropoutWrapper inside the AttentionWrapper)' \
        if graph == 'default' else 'ap4::attn_lstm_persist4_{fwd,bwd}_kernel (attention layer folded into the recurrent matrix)'
    roof = {
        'bound': 'hbm', 'achieved': round(ach, 1), 'peak': hbm_peak, 'unit': 'GB/s',
        'frac': round(ach / hbm_peak, 4),
        'traffic': dram,
        'kernel': kname + ': cross-modal audio layer T=300/Tm=75 + decoder T=41/Tm=300',
        'us_per_recurrent_step': {'fwd': round(ms['attn_lstm_fwd'] * 1e3 / steps_total, 3),
                                  'bwd': round(ms['attn_lstm_bwd'] * 1e3 / steps_total, 3),
                                  'note': 'summed kernel time / (300 + 41) steps: the layer is a latency chain '
                                          '(exchange -> product -> gate math -> sweep -> product -> exchange)'},
        'ms_per_step': {'attn_lstm_fwd': round(ms['attn_lstm_fwd'], 4), 'attn_lstm_bwd': round(ms['attn_lstm_bwd'], 4)},
        'launches_per_step': n['attn_lstm_fwd'] + n['attn_lstm_bwd'],
        'algorithmic_bytes_per_step': int(2 * bytes_dir),
        'frac_once_per_utterance': round(ach_once / hbm_peak, 5),
        'once_per_utterance_bytes_per_step': int(2 * once),
        'frac_measured_dram': round(dram / (t_att * 1e-3) / 1e9 / hbm_peak, 4) if (dram and t_att > 0) else None,
        'traffic_source': dram_src,
        'peak_source': f'hbm_gbs of {peak_src}',
        'note': '`frac` follows the contract (SURVEY.md 8d: keys + values streamed once per query step as fp32) but the '
                'kernels read fp16 copies that stay resident in the 126 MB L2, so it is an L2-fed figure against the HBM '
                'peak; by DRAM bytes (frac_measured_dram) and by the once-per-utterance bound the kernel is far from the '
                'HBM roof: it is latency bound (ncu: dram 5 %, issue slots 37 %, tensor pipe 3-4 %)',
        'timing': 'CUDA events around each launch on its stream, 3 eager steps after the graph-timed loop',
    }
    tensor = {
        'bound': 'tensor', 'unit': 'TFLOP/s',
        'recurrent_products': {
            'achieved': round(rec_tflops, 2), 'peak': f16_peak, 'frac': round(rec_tflops / f16_peak, 5),
            'kernels': 'lp4::lstm_persist4_{fwd,bwd}_kernel + the attention-LSTM kernels above (fp16 operands, fp32 '
                       'accumulation in TMEM)',
            'ms_per_step': {k: round(ms[k], 4) for k in ('lstm_fwd', 'lstm_bwd', 'attn_lstm_fwd', 'attn_lstm_bwd')},
            'flop_per_step': 2.0 * (flop_attn + flop_lstm),
            'peak_source': f'bf16 sustained of {peak_src}',
            'note': 'one 128 x 16 x 16 product chain per time step and CTA between two cluster exchanges: latency bound by '
                    'construction (sequential recurrence), the tensor pipe is ~4-6 % busy (ncu)'},
        'gemm_kernel_ms_per_step': round(ms['gemm'], 4), 'gemm_launches_per_step': n['gemm'],
        'gate_gemms': gate,
    }
    return roof, tensor


# attention layers of every BASELINE configuration: (query steps, [(memory rows, memory depth)], units)
ATT_LAYERS = {
    1: [(41, [(300, 128)], 128)],                       # LAS 1x128: decoder over the audio memory
    2: [(41, [(300, 512)], 256)],                       # BiLSTM memory (2 x 256)
    3: [(41, [(75, 256)], 256)],
    4: [(41, [(75, 256), (300, 256)], 256)],            # WLAS: video + audio mechanisms
    5: [(300, [(75, 256)], 256), (41, [(300, 256)], 256)],
}


def config_roofline(cfg, B, kms, kn):
    """The roofline object of SURVEY.md 8d-2 for one configuration's attention layers (same definition as the headline's):
    algorithmic bytes = 4 Tm (A + Dm) + 4 (Tm + Dm + A) per (utterance, query step, mechanism), both directions, over the
    persistent attention-LSTM kernels' measured time."""
    peaks = load_peaks()
    hbm_peak = float(peaks.get('hbm_gbs', 6650.0))
    t_att = kms['attn_lstm_fwd'] + kms['attn_lstm_bwd']
    if t_att <= 0:
        return {'bound': 'hbm', 'achieved': None, 'peak': hbm_peak, 'unit': 'GB/s', 'frac': None, 'traffic': None,
                'note': 'no persistent attention kernel ran: this shape (H = 128) takes the per-step launch path'}
    by, once, steps = 0.0, 0.0, 0
    for T, mems, A in ATT_LAYERS[cfg]:
        steps += T
        for Tm, Dm in mems:
            by += B * T * (4.0 * Tm * (A + Dm) + 4.0 * (Tm + Dm + A))
            once += B * 4.0 * Tm * (A + Dm)
    ach = 2.0 * by / (t_att * 1e-3) / 1e9
    return {'bound': 'hbm', 'achieved': round(ach, 1), 'peak': hbm_peak, 'unit': 'GB/s', 'frac': round(ach / hbm_peak, 4),
            'traffic': None, 'frac_once_per_utterance': round(2.0 * once / (t_att * 1e-3) / 1e9 / hbm_peak, 5),
            'us_per_recurrent_step': {'fwd': round(kms['attn_lstm_fwd'] * 1e3 / steps, 3),
                                      'bwd': round(kms['attn_lstm_bwd'] * 1e3 / steps, 3)},
            'launches_per_step': kn['attn_lstm_fwd'] + kn['attn_lstm_bwd'],
            'note': 'L2-fed figure against the HBM peak, as the headline roofline: the fp16 memories stay in the 126 MB L2'}


# ------------------------------------------------------------------------------------
# one configuration, device-resident: inputs already in HBM, K graph replays between CUDA events
# ------------------------------------------------------------------------------------
def build_model(args, torch, cfg, B, graph, seed, pinned_u8=True):
    from avsr_tf1_b200.seq2seq import Seq2SeqModel
    from tests.helpers import to_data_sequences
    hp, batch, att = workload(args, cfg, B, seed=seed, graph=graph)
    u8 = batch.pop('video_u8', None)
    pinned = {k: torch.from_numpy(v).pin_memory() for k, v in batch.items()}
    ds_float = to_data_sequences(pinned)
    ds_host = ds_float
    if u8 is not None and pinned_u8:  # the e2e loop ships the crops as stored pixels (a quarter of the bytes)
        p8 = dict(pinned)
        p8['video'] = torch.from_numpy(u8).pin_memory()
        ds_host = to_data_sequences(p8)
    model = Seq2SeqModel(ds_float, 'train', hp, seed=2001)
    model.use_cuda_graph = not args.no_graph
    model.overlap_streams = bool(args.overlap)
    return model, ds_host, att


def timed_steps(torch, model, steps, warmup, barrier=None):
    for _ in range(max(3, warmup)):
        model.train_step(fetch=False)
    (barrier or torch.cuda.synchronize)()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(steps):
        model.train_step(fetch=False)
    e1.record()
    (barrier or torch.cuda.synchronize)()
    return e0.elapsed_time(e1)


def side_config(args, torch, ops, cfg, graph, steps):
    """BASELINE config `cfg` at its own batch on one GPU: utterances/s + the kernel classes of its step."""
    B = CONFIGS[cfg]['batch']
    model, ds_host, att = build_model(args, torch, cfg, B, graph, seed=0)
    model.feed(ds_host)
    ms = timed_steps(torch, model, steps, 3) / steps
    launches = int(model.launches_last_step)
    loss, _ = model.fetch_scalars()
    kms, kn = kernel_class_times(torch, ops, model, reps=2)
    persistent = kms['attn_lstm_fwd'] + kms['attn_lstm_bwd'] + kms['lstm_fwd'] + kms['lstm_bwd']
    out = {'value': round(B / (ms / 1e3), 2), 'unit': UNIT, 'ms_per_step': round(ms, 4), 'per_gpu_batch': B,
           'gpu_launches_per_step': launches, 'loss': round(float(loss), 6), 'attention': att,
           'workload': CONFIGS[cfg]['what'], 'graph': graph,
           'kernel_ms_per_step': {k: round(v, 4) for k, v in kms.items()},
           'kernel_launches_per_step': kn,
           'persistent_kernel_share': round(persistent / ms, 3) if ms > 0 else None,
           'roofline': config_roofline(cfg, B, kms, kn)}
    del model
    torch.cuda.empty_cache()
    return out


def cnn_front_end_config(args, torch, ops, cfg, graph, steps):
    """The same workload with the reference's visual front-end in the step (video_processing='resnet_cnn', what
    run_audiovisual.py:34-58 / run_video.py select; video.py:143-195): lip crops [B, 75, 36, 36, 3] -> 13 convolutions +
    batch_norm_relu per frame -> 128-d features -> video LSTM, trained jointly.  Reports the whole step and
    the front-end's own forward + backward time."""
    from avsr_tf1_b200.seq2seq import Seq2SeqModel
    from tests.helpers import config_hparams, synthetic_batch, to_data_sequences, to_image_sequences
    B = CONFIGS[cfg]['batch']
    over = dict(randomness(graph))
    att = args.attention or ('bahdanau' if cfg == 2 else 'scaled_luong')
    over['attention_type'] = ((att,), (att,))
    hp = config_hparams(cfg, video_processing='resnet_cnn', **over)
    batch = to_image_sequences(synthetic_batch(hp, B=B, Ta=300, Tv=75, Fa=80, Fv=128, L=40, ragged=False, seed=0), hw=36)
    ds = to_data_sequences({k: torch.from_numpy(v).pin_memory() for k, v in batch.items()})
    model = Seq2SeqModel(ds, 'train', hp, seed=2001)
    model.use_cuda_graph = not args.no_graph
    model.feed(ds)
    ms = timed_steps(torch, model, steps, 3) / steps
    launches = int(model.launches_last_step)
    loss, _ = model.fetch_scalars()
    # the front-end alone, eager, CUDA events around its forward and backward
    cnn = getattr(model, '_cnn', None)
    fe = None
    if cnn is not None:
        N = B * 75
        frames = torch.rand(N, 36, 36, 3, device='cuda') * 2 - 1
        d = torch.randn(N, cnn.out_dim, device='cuda') * 1e-3
        for _ in range(2):
            cnn.forward(frames, True)
            cnn.backward(d)
        e = [torch.cuda.Event(enable_timing=True) for _ in range(3)]
        torch.cuda.synchronize()
        e[0].record()
        cnn.forward(frames, True)
        e[1].record()
        cnn.backward(d)
        e[2].record()
        torch.cuda.synchronize()
        fw, bw = e[0].elapsed_time(e[1]), e[1].elapsed_time(e[2])
        fe = {'forward_ms': round(fw, 3), 'backward_ms': round(bw, 3), 'frames': N,
              'forward_tflops': round(11.47e6 * N / fw / 1e9, 2),
              'kernels': 'cv::conv_mma_{fwd,wgrad}_kernel (implicit-GEMM mma.sync TF32, BN-ReLU fused), csrc/conv_mma.cu'}
    out = {'value': round(B / (ms / 1e3), 2), 'unit': UNIT, 'ms_per_step': round(ms, 4), 'per_gpu_batch': B,
           'gpu_launches_per_step': launches, 'loss': round(float(loss), 6), 'graph': graph,
           'video_input': '36x36x3 lip crops -> resnet_cnn (8, 16, 32, 64 filters, 128 units) inside the training step',
           'front_end': fe}
    del model
    torch.cuda.empty_cache()
    return out


def tfrecord_e2e(args, torch, model, n_utt, B):
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
            paths['video'], paths['audio'], paths['labels'], batch_size=B,
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
                'timing': 'wall clock over three epochs (each restarts the shuffling / bucketing iterator and drains the pipeline) incl. record decode, H2D (overlapped with the previous step), step, '
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
        att = args.attention or 'scaled_luong'
        hp = config_hparams(5, attention_type=((att,), (att,)), decoding_algorithm='greedy')
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

    ops.set_tensor_cores(not args.no_tensor_cores)
    cfg, graph = args.config, args.graph
    B1 = args.batch or CONFIGS[cfg]['batch']           # per-GPU batch at N = 1
    B = B1 if args.scaling == 'weak' else max(1, B1 // world)

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

    model, ds_host, att = build_model(args, torch, cfg, B, graph, seed=rank)

    # ---- device-resident timing: inputs already in HBM ------------------------------------
    model.feed(ds_host)
    sampler = ClockSampler(local)
    sampler.start()
    ms_total = max_over_ranks(timed_steps(torch, model, args.steps, args.warmup, barrier))
    launches = model.launches_last_step
    loss, gnorm = model.fetch_scalars()

    # ---- end to end: pinned host batch -> H2D -> step -> D2H loss, every step ---------------
    # The H2D copy of step k+1's batch is issued (copy stream) right after step k is launched, so it overlaps
    # the compute of step k - the input-pipeline prefetch the reference gets from tf.data (io_utils.py:145).
    # Every step still copies its full batch from pinned host memory (lip crops as the stored uint8 pixels, expanded
    # to the reference's floats on the device) and reads its loss back (the read of step k is waited for after step
    # k+1 has been launched, so the host round trip is off the GPU's critical path).
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

    # ---- N > 1: the other scaling mode beside the headline (global batch 256 at 1/2/4/8 GPUs) ----
    other = None
    if world > 1 and not args.skip_extras:
        try:
            Bo = max(1, B1 // world) if args.scaling == 'weak' else B1
            del model
            torch.cuda.empty_cache()
            model, ds_o, _ = build_model(args, torch, cfg, Bo, graph, seed=rank)
            model.feed(ds_o)
            ms_o = max_over_ranks(timed_steps(torch, model, args.steps, args.warmup, barrier)) / args.steps
            other = {'scaling': 'strong' if args.scaling == 'weak' else 'weak', 'per_gpu_batch': Bo,
                     'global_batch': Bo * world, 'ms_per_step': round(ms_o, 4),
                     'value': round(Bo * world / (ms_o / 1e3), 2), 'unit': UNIT,
                     'note': 'device-timed like `value`; strong scaling of a latency-bound recurrence: the step time barely '
                             'falls with the per-GPU batch (one wave of 8-utterance clusters either way)'}
        except Exception as ex:
            other = {'error': repr(ex)}

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
    value = B * world / (ms_step / 1e3)
    e2e_step = e2e_ms / args.steps
    line = {
        'metric': METRICS[cfg], 'value': round(value, 2), 'unit': UNIT, 'n_gpus': world, 'steps': args.steps,
        'warmup': max(3, args.warmup), 'ms_per_step': round(ms_step, 4), 'higher_is_better': True,
        'scaling': args.scaling, 'vs_baseline': None,
        'dtype': 'f32' if args.no_tensor_cores else 'f32 storage/accumulate, tf32 tensor-core products',
        'data': 'synthetic', 'config': config_dict(args, cfg, B, world, att, graph, args.scaling),
        'clocks': sampler.summary(),
        'e2e': {'value': round(B * world / (e2e_step / 1e3), 2), 'unit': UNIT, 'ms_per_step': round(e2e_step, 4),
                'wall_ms_per_step': round(e2e_wall / args.steps, 4), 'h2d_bytes_per_step': int(h2d),
                'd2h_bytes_per_step': d2h},
        'gpu_launches': int(launches * args.steps),
        'gpu_launches_per_step': int(launches),
        'cuda_graph': not args.no_graph,
        'loss': round(float(loss), 6), 'global_norm': round(float(gnorm), 6), 'n_params': int(model.n_params),
    }
    if other is not None:
        line['strong_scaling' if args.scaling == 'weak' else 'weak_scaling'] = other
    try:
        if args.skip_roofline or world > 1:
            line['roofline'] = None
        elif cfg != 5:  # --config 1..4 as the headline: the same object as in the `configs` side key of the default run
            kms, kn = kernel_class_times(torch, ops, model)
            line['roofline'] = config_roofline(cfg, B, kms, kn)
            line['kernel_ms_per_step'] = {k: round(v, 4) for k, v in kms.items()}
        else:
            gate = gate_gemm_roofline(args, torch, ops, B)
            kms, kn = kernel_class_times(torch, ops, model)
            line['roofline'], line['roofline_tensor'] = attention_roofline(args, B, graph, kms, kn, gate)
    except Exception as ex:  # keep the headline number even if the side measurement fails
        line['roofline'] = {'error': repr(ex)}
    if world == 1 and not args.skip_extras:
        if cfg == 5:
            try:
                line['e2e_tfrecord'] = tfrecord_e2e(args, torch, model, args.tfrecord_utterances, B)
            except Exception as ex:
                line['e2e_tfrecord'] = {'error': repr(ex)}
        del model
        torch.cuda.empty_cache()
        side_steps = max(2, min(args.steps, 5))
        if graph == 'default':  # the same workload with the randomness off: the graph the parity tests check
            try:
                line['parity_graph'] = side_config(args, torch, ops, cfg, 'parity', side_steps)
            except Exception as ex:
                line['parity_graph'] = {'error': repr(ex)}
        if cfg >= 3:  # the visual front-end inside the step (what run_video.py / run_audiovisual.py train)
            try:
                line['with_resnet_cnn'] = cnn_front_end_config(args, torch, ops, cfg, graph, side_steps)
            except Exception as ex:
                line['with_resnet_cnn'] = {'error': repr(ex)}
        line['configs'] = {}
        for c in sorted(CONFIGS):  # the other BASELINE.json configurations, each at its own batch, same graph
            if c == cfg:
                continue
            try:
                line['configs'][str(c)] = side_config(args, torch, ops, c, graph, side_steps)
            except Exception as ex:
                line['configs'][str(c)] = {'error': repr(ex)}
    if world == 1 and not args.skip_cpu_baseline:
        r = run_oracle(args, cfg, graph, steps=3, warmup=1, sample=min(args.cpu_sample, B))
        line['cpu_baseline'] = {
            'value': round(r['value'], 3), 'unit': UNIT, 'cores': r['cores'], 'kind': 'port',
            'sample': f'{r["sample"]} utterances per step of the same workload and graph at full sequence lengths, 3 timed '
                      'steps (about 10 s of host work); NumPy/OpenBLAS restatement of the TF1 graph (oracle/)'}
        # matrix-product FLOP of one training step per utterance (forward x 3: forward, input gradient, weight gradient)
        # -> the rate the CPU arm sustains, so its distance from the host's own roof can be judged
        fl = train_flop_per_utterance(cfg, 3888 if args.video_input == 'crops3888' else 128)
        line['cpu_baseline']['gflops'] = round(r['value'] * fl / 1e9, 1)
        line['cpu_baseline']['gflop_per_utterance'] = round(fl / 1e9, 2)
        if cfg != 1:  # BASELINE.json configs[0]: the reference's own CPU-runnable case, at its own batch of 2
            try:
                r1 = run_oracle(args, 1, graph, steps=3, warmup=1, sample=CONFIGS[1]['batch'])
                line['cpu_baseline']['config_1'] = {
                    'value': round(r1['value'], 3), 'unit': UNIT, 'cores': r1['cores'], 'ms_per_step': round(r1['sec_per_step'] * 1e3, 1),
                    'sample': 'configs[0] exactly: audio-only LAS 1x128, batch 2, Ta = 300 x 80, 41 label steps, 3 timed steps'}
            except Exception as ex:
                line['cpu_baseline']['config_1'] = {'error': repr(ex)}
        try:  # in a child process: nothing it does can cost the parent its JSON line
            child = subprocess.run([sys.executable, os.path.abspath(__file__), '--cer-check-only'] +
                                   (['--attention', args.attention] if args.attention else []),
                                   capture_output=True, text=True, timeout=300)
            out_lines = [ln for ln in child.stdout.strip().splitlines() if ln.startswith('{')]
            line['cpu_baseline']['cer_check'] = json.loads(out_lines[-1]) if out_lines else {
                'error': 'exit %d: %s' % (child.returncode, child.stderr.strip()[-300:])}
        except Exception as ex:
            line['cpu_baseline']['cer_check'] = {'error': repr(ex)}
    print(json.dumps(line), flush=True)
    finish()


if __name__ == '__main__':
    main()
