"""The C-ABI library loads and exports exactly the symbols include/avsr_b200.h declares, and the ctypes
prototypes cover all of them (no compute calls: runs without a GPU)."""
import ctypes
import os
import re
import subprocess

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
HEADER = os.path.join(ROOT, 'include', 'avsr_b200.h')


def declared_symbols():
    text = open(HEADER).read()
    text = re.sub(r'/\*.*?\*/', '', text, flags=re.S)
    return sorted(set(re.findall(r'\b(avsr_[a-z0-9_]+)\s*\(', text)))


@pytest.fixture(scope='module')
def lib():
    from avsr_tf1_b200 import _lib
    if not os.path.exists(_lib.LIB_PATH):
        subprocess.run(['make', '-j8', '-C', ROOT], check=True)
    return _lib


def test_header_symbols_are_exported(lib):
    syms = declared_symbols()
    assert len(syms) >= 25
    cdll = ctypes.CDLL(lib.LIB_PATH)
    for s in syms:
        assert hasattr(cdll, s), f'{s} declared in include/avsr_b200.h but not exported'


def test_prototypes_cover_header_both_ways(lib):
    assert sorted(lib.PROTOTYPES) == declared_symbols()


def test_library_loads_and_reports_version(lib):
    l = lib.load()
    assert l.avsr_version() >= 100
    assert l.avsr_last_error() is not None
    assert l.avsr_launch_count() == 0 or l.avsr_launch_count() > 0
    old = l.avsr_set_tensor_cores(0)
    assert l.avsr_get_tensor_cores() == 0
    l.avsr_set_tensor_cores(old)


def test_struct_sizes_match_header(lib):
    sizes = (ctypes.c_int * 2)()
    lib.load().avsr_struct_sizes(sizes)  # sizeof() as compiled by nvcc
    assert ctypes.sizeof(lib.AvsrAttnMech) == sizes[0] == 16 + 22 * 8
    assert ctypes.sizeof(lib.AvsrRnnSeq) == sizes[1]


def test_no_cpu_fallback_when_library_missing(lib, monkeypatch):
    monkeypatch.setattr(lib, '_lib', None)
    monkeypatch.setattr(lib, 'LIB_PATH', '/nonexistent/libavsr_b200.so')
    with pytest.raises(lib.AvsrError):
        lib.load()
