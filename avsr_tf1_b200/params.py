"""Parameter table: one flat fp32 device buffer (+ gradient, Adam m / v) with a
name -> (offset, shape) table keyed by the TF variable names the reference graph
would create, so checkpoints can be exchanged by name."""
from __future__ import annotations

import math
import zlib
from dataclasses import dataclass
from typing import Dict, List, Tuple

import numpy as np
import torch

ALIGN = 64  # floats (256 B): keeps every tensor TMA / float4 aligned


@dataclass
class ParamSpec:
    name: str
    shape: Tuple[int, ...]
    init: str  # lstm_kernel | conv_kernel | glorot | embedding | zeros | ones | const:<v>
    trainable: bool = True


def _truncated_normal(rng, shape, std):
    out = rng.standard_normal(shape)
    bad = np.abs(out) > 2.0
    while bad.any():
        out[bad] = rng.standard_normal(int(bad.sum()))
        bad = np.abs(out) > 2.0
    return out * std


def init_array(spec: ParamSpec, seed: int, vocab: int = 31) -> np.ndarray:
    """TF-1.13 style initialisers restated (cells.py:17 variance_scaling; Dense default
    glorot_uniform; decoder_unimodal.py:80-83 embedding)."""
    rng = np.random.default_rng([seed, zlib.crc32(spec.name.encode())])
    shape = spec.shape
    if spec.init == 'zeros':
        a = np.zeros(shape)
    elif spec.init == 'ones':
        a = np.ones(shape)
    elif spec.init.startswith('const:'):
        a = np.full(shape, float(spec.init.split(':')[1]))
    elif spec.init == 'lstm_kernel':
        fan_in = shape[0]
        a = _truncated_normal(rng, shape, math.sqrt(1.0 / fan_in) / .87962566103423978)
    elif spec.init == 'conv_kernel':  # video.py:26 variance_scaling_initializer(scale=2.0, mode='fan_in')
        fan_in = int(np.prod(shape[:-1]))
        a = _truncated_normal(rng, shape, math.sqrt(2.0 / fan_in) / .87962566103423978)
    elif spec.init == 'glorot':
        fan_in, fan_out = (shape[0], shape[1]) if len(shape) == 2 else (shape[0], shape[0])
        lim = math.sqrt(6.0 / (fan_in + fan_out))
        a = rng.uniform(-lim, lim, shape)
    elif spec.init == 'embedding':
        lim = 1.732 / vocab
        a = rng.uniform(-lim, lim, shape)
    else:
        raise ValueError('unknown initialiser ' + spec.init)
    return np.asarray(a, np.float32).reshape(shape)


class ParamStore:
    def __init__(self, specs: List[ParamSpec], device='cuda', with_optimizer=True):
        self.specs = specs
        self.table: Dict[str, Tuple[int, Tuple[int, ...]]] = {}
        off = 0
        for s in specs:
            if not s.trainable:
                continue
            n = int(np.prod(s.shape)) if len(s.shape) else 1
            self.table[s.name] = (off, tuple(s.shape))
            off += (n + ALIGN - 1) // ALIGN * ALIGN
        self.numel = off
        self.n_trainable = sum(int(np.prod(s.shape)) if len(s.shape) else 1 for s in specs if s.trainable)
        self.flat = torch.zeros(max(off, ALIGN), dtype=torch.float32, device=device)
        # tf32-rounded copy of the parameters: the operand the tcgen05 products read (kept in sync by the
        # Adam kernel; refreshed by sync_tf32() after initialisation / restore)
        self.flat_tc = torch.zeros_like(self.flat)
        self.grad = torch.zeros_like(self.flat) if with_optimizer else None
        self.m = torch.zeros_like(self.flat) if with_optimizer else None
        self.v = torch.zeros_like(self.flat) if with_optimizer else None
        self.state: Dict[str, torch.Tensor] = {
            s.name: torch.zeros(s.shape, dtype=torch.float32, device=device) for s in specs if not s.trainable}

    def _view(self, buf, name):
        off, shape = self.table[name]
        n = int(np.prod(shape)) if len(shape) else 1
        return buf[off:off + n].view(shape if len(shape) else (1,))

    def p(self, name) -> torch.Tensor:
        return self.state[name] if name in self.state else self._view(self.flat, name)

    def w(self, name, rounded: bool) -> torch.Tensor:
        """Parameter as a matrix-product operand: the tf32-rounded copy in tensor-core mode."""
        if name in self.state or not rounded:
            return self.p(name)
        return self._view(self.flat_tc, name)

    def sync_tf32(self):
        if self.flat.is_cuda:
            from . import ops
            ops.round_tf32(self.flat, self.flat_tc)

    def g(self, name) -> torch.Tensor:
        return self._view(self.grad, name)

    def names(self, trainable_only=True):
        return [s.name for s in self.specs if s.trainable or not trainable_only]

    def initialize(self, seed: int, vocab: int = 31):
        for s in self.specs:
            self.p(s.name).copy_(torch.from_numpy(init_array(s, seed, vocab)).view(self.p(s.name).shape))
        self.sync_tf32()

    def load_numpy(self, arrays: Dict[str, np.ndarray], strict=True):
        for s in self.specs:
            if s.name in arrays:
                a = np.asarray(arrays[s.name], np.float32)
                self.p(s.name).copy_(torch.from_numpy(a).reshape(self.p(s.name).shape))
            elif strict:
                raise KeyError('missing variable ' + s.name)
        self.sync_tf32()

    def to_numpy(self, what='p') -> Dict[str, np.ndarray]:
        out = {}
        for s in self.specs:
            if what == 'p':
                t = self.p(s.name)
            elif s.trainable:
                t = self._view({'g': self.grad, 'm': self.m, 'v': self.v}[what], s.name)
            else:
                continue
            out[s.name] = t.detach().cpu().numpy().reshape(s.shape).copy()
        return out
