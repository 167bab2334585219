"""The boundary is a C ABI: a plain C99 program (tests/c/abi_smoke.c) includes include/avsr_b200.h, links against
libavsr_b200.so and the CUDA runtime only, and drives the path without Python or torch.  CPU: it compiles and links
(the header is valid C, every symbol it uses resolves).  GPU: it runs and checks its results against its own loops."""
import os
import shutil
import subprocess

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
CUDA = os.environ.get('CUDA_HOME', '/usr/local/cuda')
LIBDIR = os.path.join(ROOT, 'avsr_tf1_b200', 'lib')


def build(out):
    cc = shutil.which('gcc') or shutil.which('cc')
    if cc is None or not os.path.exists(os.path.join(CUDA, 'include', 'cuda_runtime_api.h')):
        pytest.skip('no C compiler / CUDA headers')
    cmd = [cc, '-std=c99', '-O1', '-Wall', '-Werror', os.path.join(ROOT, 'tests', 'c', 'abi_smoke.c'),
           '-I' + os.path.join(ROOT, 'include'), '-I' + os.path.join(CUDA, 'include'), '-L' + LIBDIR, '-lavsr_b200',
           '-L' + os.path.join(CUDA, 'lib64'), '-lcudart', '-lm', '-o', out]
    r = subprocess.run(cmd, capture_output=True, text=True)
    assert r.returncode == 0, r.stderr
    return out


def test_c_client_compiles_and_links(tmp_path):
    build(str(tmp_path / 'abi_smoke'))


@pytest.mark.gpu
def test_c_client_runs_on_the_gpu(tmp_path):
    exe = build(str(tmp_path / 'abi_smoke'))
    env = dict(os.environ)
    env['LD_LIBRARY_PATH'] = os.pathsep.join([LIBDIR, os.path.join(CUDA, 'lib64'), env.get('LD_LIBRARY_PATH', '')])
    r = subprocess.run([exe], capture_output=True, text=True, env=env, timeout=120)
    assert r.returncode == 0, r.stdout + r.stderr
    assert r.stdout.strip().endswith('OK'), r.stdout
