"""Generator / Discriminator with the stand-in's module API (oracle/cyclegan_standin.py:129,172):
`forward(x)`, `state_dict()`, `load_state_dict()`, `named_parameters()`, `parameters()`.
Parameters are fp32 torch tensors; once a module is attached to a step engine they become views of
the engine's flat parameter buffer, so training is visible through `state_dict()`.
All compute runs in libcyclegan_b200.so."""
from __future__ import annotations

from collections import OrderedDict
from typing import Dict, Iterator, Optional, Tuple

import torch

from . import engine as _engine


class _NetModule:
    _KIND = "G"

    def __init__(self, n_blocks: int = 9, seed: Optional[int] = None, device=None):
        self.n_blocks = n_blocks
        inv = _engine.describe(1, 64, n_blocks)
        self._infos = inv[0 if self._KIND == "G" else 2]
        self._device = torch.device(device) if device is not None else (
            torch.device("cuda", torch.cuda.current_device()) if torch.cuda.is_available() else torch.device("cpu"))
        g = torch.Generator().manual_seed(seed) if seed is not None else None
        self._params: "OrderedDict[str, torch.Tensor]" = OrderedDict()
        for info in self._infos:
            if info.is_bias:
                t = torch.zeros(info.torch_shape)
            else:
                t = torch.empty(info.torch_shape).normal_(0.0, 0.02, generator=g)
            self._params[info.name] = t.to(self._device)
        self._engine: Optional[_engine.StepEngine] = None  # set by CycleGANTrainer
        self._net: Optional[int] = None
        # inference-only engines for input shapes other than the trainer's, newest last; each remembers the weight
        # version it was packed from and is re-synchronised when the weights moved on (training steps, in-place writes)
        self._private: "OrderedDict[Tuple[int, int], _engine.StepEngine]" = OrderedDict()
        self._private_version: Dict[Tuple[int, int], Tuple[int, int]] = {}
        self._trained_steps = 0  # bumped by CycleGANTrainer after every optimiser step (raw-pointer writes)

    _MAX_PRIVATE = 2  # private inference engines kept alive per module (each holds one forward pass of workspace)

    def _weights_version(self) -> Tuple[int, int]:
        """changes whenever the parameters may have changed: optimiser steps of the attached trainer, plus torch's
        in-place version counters (writes through parameters() / named_parameters() views)"""
        return self._trained_steps, sum(int(p._version) for p in self._params.values())

    # ---- nn.Module-like surface -------------------------------------------------------------------
    def named_parameters(self) -> Iterator[Tuple[str, torch.Tensor]]:
        return iter(self._params.items())

    def parameters(self) -> Iterator[torch.Tensor]:
        return iter(self._params.values())

    def state_dict(self) -> "OrderedDict[str, torch.Tensor]":
        return OrderedDict((k, v.detach().clone()) for k, v in self._params.items())

    def load_state_dict(self, sd) -> None:
        missing = [k for k in self._params if k not in sd]
        extra = [k for k in sd if k not in self._params]
        if missing or extra:
            raise KeyError(f"state_dict mismatch: missing {missing}, unexpected {extra}")
        for k, dst in self._params.items():
            src = sd[k]
            if tuple(src.shape) != tuple(dst.shape):
                raise ValueError(f"{k}: shape {tuple(src.shape)} != {tuple(dst.shape)}")
            dst.copy_(src.to(device=dst.device, dtype=torch.float32))
        if self._engine is not None:
            self._engine.refresh_weights(0 if self._KIND == "G" else 1)

    # ---- engine plumbing --------------------------------------------------------------------------
    def _attach(self, eng: _engine.StepEngine, net: int) -> None:
        views = eng.param_views(net)
        for k, v in views.items():
            v.copy_(self._params[k].to(v.device))
            self._params[k] = v
        self._engine, self._net = eng, net
        self._device = eng.device
        self._private.clear()
        self._private_version.clear()

    def _engine_for(self, x: torch.Tensor):
        batch, _, h, w = x.shape
        if h != w:
            raise ValueError("only square images are supported")
        if self._engine is not None and (self._engine.batch, self._engine.size) == (batch, h):
            return self._engine, self._net
        key = (batch, h)
        net = 0 if self._KIND == "G" else 2
        if key not in self._private:
            while len(self._private) >= self._MAX_PRIVATE:  # evict the least recently used engine
                old, _ = self._private.popitem(last=False)
                self._private_version.pop(old, None)
            self._private[key] = _engine.StepEngine(batch, h, self.n_blocks, inference=True)
        self._private.move_to_end(key)
        eng = self._private[key]
        version = self._weights_version()
        if self._private_version.get(key) != version:  # first use, or the weights changed since it was packed
            for k, v in eng.param_views(net).items():
                v.copy_(self._params[k].to(v.device))
            eng.refresh_weights(0 if self._KIND == "G" else 1)
            self._private_version[key] = self._weights_version()
        return eng, net

    def __call__(self, x: torch.Tensor) -> torch.Tensor:
        return self.forward(x)


class Generator(_NetModule):
    """ResNet-9-block generator: forward(x:[N,3,H,H] in [-1,1]) -> [N,3,H,H]."""
    _KIND = "G"

    def __init__(self, in_ch: int = 3, out_ch: int = 3, ngf: int = 64, n_blocks: int = 9, seed=None, device=None):
        if (in_ch, out_ch, ngf) != (3, 3, 64):
            raise ValueError("the B200 path pins the canonical CycleGAN generator (3 -> 3 channels, ngf=64)")
        super().__init__(n_blocks, seed, device)

    def forward(self, x: torch.Tensor) -> torch.Tensor:
        eng, net = self._engine_for(x)
        return eng.generator_forward(net, x)


class Discriminator(_NetModule):
    """70x70 PatchGAN: forward(x:[N,3,H,H]) -> [N,1,H/8-2,H/8-2]."""
    _KIND = "D"

    def __init__(self, in_ch: int = 3, ndf: int = 64, n_layers: int = 3, seed=None, device=None):
        if (in_ch, ndf, n_layers) != (3, 64, 3):
            raise ValueError("the B200 path pins the canonical 70x70 PatchGAN (3 channels, ndf=64, n_layers=3)")
        super().__init__(9, seed, device)

    def forward(self, x: torch.Tensor) -> torch.Tensor:
        eng, net = self._engine_for(x)
        return eng.discriminator_forward(net, x)
