"""bench.py's measurement contract, exercised for real on the CPU arm (the GPU arm needs a B200): the reference arm
of the smallest BASELINE configuration runs in a child process launched the way torchrun launches workers
(OMP_NUM_THREADS=1) and must print ONE JSON line carrying the keys the driver reads, on the same graph and
configuration names the product arm reports, using more than the one thread the launcher allowed."""
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def run_reference(*extra):
    env = dict(os.environ, OMP_NUM_THREADS='1')
    out = subprocess.run([sys.executable, os.path.join(ROOT, 'bench.py'), '--impl', 'reference', '--config', '1',
                          '--steps', '2', '--warmup', '1', '--cpu-sample', '2', *extra],
                         capture_output=True, text=True, timeout=600, env=env, cwd=ROOT)
    assert out.returncode == 0, out.stderr[-2000:]
    lines = [ln for ln in out.stdout.splitlines() if ln.strip()]
    assert len(lines) == 1, lines  # exactly one JSON line
    return json.loads(lines[0])


def test_reference_arm_line_default_graph():
    d = run_reference()
    for k in ('impl', 'metric', 'value', 'unit', 'n_gpus', 'steps', 'warmup', 'ms_per_step', 'higher_is_better',
              'scaling', 'vs_baseline', 'dtype', 'data', 'config', 'cpu_baseline', 'e2e'):
        assert k in d, k
    assert d['impl'] == 'reference' and d['unit'] == 'utterances/s' and d['higher_is_better'] is True
    assert d['steps'] == 2 and d['warmup'] == 1  # the arm runs the K / W it was given (no clamping)
    assert d['value'] > 0 and abs(d['value'] - d['cpu_baseline']['value']) < 1e-9
    assert d['e2e'] == {'value': d['value'], 'unit': d['unit'], 'h2d_bytes_per_step': 0, 'd2h_bytes_per_step': 0}
    assert d['cpu_baseline']['kind'] == 'port'
    assert d['cpu_baseline']['cores'] == len(os.sched_getaffinity(0))  # the launcher's OMP_NUM_THREADS=1 is overridden
    assert 'configs[0]' in d['config']['workload']
    assert d['config']['dropout'].startswith('DropoutWrapper') and d['config']['scheduled_sampling'].startswith('0.1')
    assert abs(d['ms_per_step'] * d['value'] / 1e3 - 2.0) < 1e-6  # 2 utterances per step


def test_reference_arm_parity_graph_and_nonzero_rank():
    d = run_reference('--graph', 'parity')
    assert d['config']['dropout'] == 'off' and d['config']['scheduled_sampling'] == 'off'
    # under torchrun only rank 0 runs the CPU arm; the other ranks exit 0 without output
    env = dict(os.environ, RANK='1', WORLD_SIZE='2')
    out = subprocess.run([sys.executable, os.path.join(ROOT, 'bench.py'), '--impl', 'reference', '--config', '1',
                          '--gpus', '2'], capture_output=True, text=True, timeout=120, env=env, cwd=ROOT)
    assert out.returncode == 0 and out.stdout.strip() == ''


def test_roofline_arithmetic_matches_survey_8d():
    """The algorithmic bytes / FLOP that `roofline` divides by are SURVEY.md 8d's per-unit figures times the units of one
    step (no GPU needed: the kernel times are given)."""
    sys.path.insert(0, ROOT)
    import bench

    class Args:
        video_input = 'crops3888'
    ms = {'attn_lstm_fwd': 3.0, 'attn_lstm_bwd': 3.0, 'lstm_fwd': 2.0, 'lstm_bwd': 1.0, 'gemm': 2.5}
    n = {'attn_lstm_fwd': 2, 'attn_lstm_bwd': 2, 'lstm_fwd': 5, 'lstm_bwd': 5, 'gemm': 39}
    gate = {'frac': 0.5, 'achieved': 350.0, 'peak': 700.0}
    roof, tensor = bench.attention_roofline(Args(), 256, 'default', ms, n, gate)
    # 4 Tm (A + Dm) + 4 (Tm + Dm + A) bytes per (utterance, query step): 155 948 B at Tm = 75, 617 648 B at Tm = 300
    per_dir = 256 * (300 * 155948 + 41 * 617648)
    assert roof['algorithmic_bytes_per_step'] == 2 * per_dir
    assert roof['once_per_utterance_bytes_per_step'] == 2 * 256 * 4 * (75 + 300) * 512
    assert abs(roof['achieved'] - 2 * per_dir / 6e-3 / 1e9) < 1e-3 * roof['achieved']
    assert abs(roof['frac'] - roof['achieved'] / roof['peak']) < 1e-3
    assert roof['bound'] == 'hbm' and roof['unit'] == 'GB/s'
    # matrix-product FLOP of a training step per utterance: dominated by 3 x 2 (I + H) 4H per (layer, step)
    f5 = bench.train_flop_per_utterance(5, 3888)
    lower = 3 * 2 * 4 * 256 * (75 * (3888 + 256) + 2 * 75 * 512 + 300 * (80 + 256) + 300 * 512 + 300 * 768 + 41 * (128 + 512))
    assert lower <= f5 <= 1.1 * lower
    assert bench.train_flop_per_utterance(1, 128) < 0.1 * f5
