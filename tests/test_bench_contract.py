"""The JSON line bench.py printed on the B200 (committed under profiles/) carries every key of the measurement contract,
for both arms, and the numbers are mutually consistent."""
import json
import os

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def load(name):
    with open(os.path.join(ROOT, 'profiles', name)) as f:
        return json.loads(f.read().strip().splitlines()[-1])


def test_b200_arm_line():
    l = load('r01_bench_v23_final.json')
    for k in ('metric', 'value', 'unit', 'n_gpus', 'steps', 'warmup', 'ms_per_step', 'higher_is_better', 'scaling',
              'vs_baseline', 'dtype', 'data', 'config', 'clocks', 'e2e', 'gpu_launches', 'roofline', 'cpu_baseline'):
        assert k in l, k
    assert l['metric'].startswith('AV-Align train utterances/sec') and l['unit'] == 'utterances/s'
    assert l['n_gpus'] == 1 and l['warmup'] >= 3 and l['higher_is_better'] is True and l['scaling'] == 'weak'
    assert l['vs_baseline'] is None and l['data'] == 'synthetic' and 'workload' in l['config']
    assert 'model' not in l['config']
    # value = utterances processed / device time
    assert abs(l['value'] - l['config']['global_batch'] / (l['ms_per_step'] / 1e3)) <= 1e-3 * l['value']
    e = l['e2e']
    assert e['unit'] == l['unit'] and e['h2d_bytes_per_step'] > 3e8 and e['d2h_bytes_per_step'] > 0
    assert 0.5 * l['value'] < e['value'] < l['value']  # end to end is measured, not copied
    assert l['gpu_launches'] == l['gpu_launches_per_step'] * l['steps'] > 0
    c = l['clocks']
    assert c['sm_mhz'] > 0.9 * c['sm_max_mhz'] and not set(c['reasons']) & {'hw_slowdown', 'hw_thermal_slowdown',
                                                                           'sw_thermal_slowdown'}
    r = l['roofline']
    assert r['bound'] in ('hbm', 'tensor') and r['unit'] in ('GB/s', 'TFLOP/s')
    assert abs(r['frac'] - r['achieved'] / r['peak']) < 1e-3 and r['traffic'] is not None
    b = l['cpu_baseline']
    assert b['kind'] in ('port', 'reference') and b['cores'] >= 1 and b['unit'] == l['unit'] and b['sample']
    assert l['value'] > 1000 * b['value']


def test_reference_arm_line():
    l = load('r01_bench_v23_reference_arm.json')
    assert l['impl'] == 'reference'
    b200 = load('r01_bench_v23_final.json')
    for k in ('metric', 'unit', 'higher_is_better'):
        assert l[k] == b200[k]
    assert l['config']['workload'] == b200['config']['workload']
    assert l['e2e'] == {'value': l['value'], 'unit': l['unit'], 'h2d_bytes_per_step': 0, 'd2h_bytes_per_step': 0}
    assert l['cpu_baseline']['value'] == l['value'] and l['cpu_baseline']['kind'] == 'port'
