"""CycleGANTrainer with the stand-in's API (oracle/cyclegan_standin.py:278): same constructor
arguments, `train_step(real_A, real_B) -> dict`, `forward_only`, `backward_only`.  One process per
GPU; with torch.distributed initialised the two flat gradient buffers are all-reduced on a side
stream, overlapped with the discriminator phase."""
from __future__ import annotations

import os

from collections import OrderedDict
from typing import Dict, Optional

import torch

from . import engine as _engine
from .modules import Discriminator, Generator
from .parallel import GradSync


class _PoolDecisions:
    """Host side of the image history pool: the random decisions of the canonical ImagePool.query, one (store, ret)
    pair per image (restated from the stand-in's PoolDecisions, oracle/cyclegan_standin.py; the images themselves
    live in the engine's workspace and are exchanged by a kernel)."""

    def __init__(self, pool_size: int, seed: int):
        import random
        self.size, self.count, self.rng = int(pool_size), 0, random.Random(seed)

    def next(self):
        if self.size <= 0:
            return -1, -1
        if self.count < self.size:
            self.count += 1
            return self.count - 1, -1
        if self.rng.uniform(0.0, 1.0) > 0.5:
            i = self.rng.randint(0, self.size - 1)
            return i, i
        return -1, -1


class CycleGANTrainer:
    LOSS_KEYS = _engine.LOSS_KEYS

    def __init__(self, G_AB: Generator, G_BA: Generator, D_A: Discriminator, D_B: Discriminator,
                 lambda_A: float = 10.0, lambda_B: float = 10.0, lambda_idt: float = 0.5, lr: float = 2e-4,
                 betas=(0.5, 0.999), eps: float = 1e-8, process_group=None, pool_size: int = 0, pool_seed: int = 0):
        self.G_AB, self.G_BA, self.D_A, self.D_B = G_AB, G_BA, D_A, D_B
        if G_AB.n_blocks != G_BA.n_blocks:
            raise ValueError("both generators must have the same number of residual blocks")
        self._hyper = dict(lambda_A=lambda_A, lambda_B=lambda_B, lambda_idt=lambda_idt, lr=lr, betas=betas, eps=eps)
        self.sync = GradSync(process_group)
        self.pool_size = int(pool_size)
        self._pools = (_PoolDecisions(pool_size, pool_seed), _PoolDecisions(pool_size, pool_seed + 1))  # fake_B, fake_A
        self.engine: Optional[_engine.StepEngine] = None
        self.stream: Optional[torch.cuda.Stream] = None
        self.comm_stream: Optional[torch.cuda.Stream] = None
        self._dp_segmented = bool(os.environ.get("CGB_DP_SEGMENTED"))  # the round-1 segmented data-parallel step

    # ---- engine --------------------------------------------------------------------------------------
    def _ensure_engine(self, real_A: torch.Tensor) -> _engine.StepEngine:
        batch, _, h, w = real_A.shape
        if h != w:
            raise ValueError("only square images are supported")
        if self.engine is not None:
            if (self.engine.batch, self.engine.size) != (batch, h):
                raise ValueError("the trainer is bound to input shape "
                                 f"{(self.engine.batch, 3, self.engine.size, self.engine.size)}")
            return self.engine
        eng = _engine.StepEngine(batch, h, self.G_AB.n_blocks, pool_size=self.pool_size, **self._hyper)
        for net, mod in enumerate((self.G_AB, self.G_BA, self.D_A, self.D_B)):
            mod._attach(eng, net)
        eng.refresh_weights(0)
        eng.refresh_weights(1)
        eng.set_grad_scale(self.sync.grad_scale)
        self.engine = eng
        self.stream = torch.cuda.Stream(device=eng.device)
        self.comm_stream = torch.cuda.Stream(device=eng.device)
        if getattr(self, "_lr", None) is not None:  # set_lr() before the first step
            eng.set_lr(self._lr)
        return eng

    def _pool_step(self, eng) -> None:
        """this step's image-pool decisions (one per image and pool), sent ahead of the D phase"""
        if self.pool_size <= 0:
            return
        dec = torch.empty(2, eng.batch, 2, dtype=torch.int32)
        for side in range(2):
            for n in range(eng.batch):
                store, ret = self._pools[side].next()
                dec[side, n, 0], dec[side, n, 1] = store, ret
        dec = dec.pin_memory()
        with torch.cuda.stream(self.stream):
            eng.set_pool_decisions(dec)

    def _enter(self):
        self.stream.wait_stream(torch.cuda.current_stream())
        return torch.cuda.stream(self.stream)

    def _exit(self):
        torch.cuda.current_stream().wait_stream(self.stream)

    # ---- API -----------------------------------------------------------------------------------------
    def forward_only(self, real_A: torch.Tensor, real_B: torch.Tensor) -> "OrderedDict[str, torch.Tensor]":
        eng = self._ensure_engine(real_A)
        with self._enter():
            eng.set_inputs(real_A, real_B)
            eng.forward_cycle()
            out = OrderedDict((k, eng.get_image(k)) for k in ("fake_B", "rec_A", "fake_A", "rec_B", "idt_A", "idt_B"))
        self._exit()
        return out

    def backward_only(self, real_A: torch.Tensor, real_B: torch.Tensor) -> Dict[str, float]:
        """forward + both backward phases, no optimiser step; gradients are left in `grads(name)`"""
        eng = self._ensure_engine(real_A)
        self._pool_step(eng)
        with self._enter():
            eng.set_inputs(real_A, real_B)
            eng.phase_generators()
            eng.phase_discriminators()
            losses = eng.losses()
        self._exit()
        return losses

    def set_lr(self, lr: float) -> None:
        """learning rate of both optimisers from the next step on (stand-in: CycleGANTrainer.set_lr)"""
        self._lr = float(lr)
        if self.engine is not None:
            with self._enter():
                self.engine.set_lr(self._lr)
            self._exit()

    def train_step(self, real_A: torch.Tensor, real_B: torch.Tensor) -> Dict[str, float]:
        if real_A.dtype == torch.uint8:
            return self._train_step_u8(real_A, real_B)
        eng = self._ensure_engine(real_A)
        self._pool_step(eng)
        if self.sync.world_size == 1:
            if real_A.device.type == "cpu":
                with self._enter():
                    losses = eng.train_step_host(real_A.contiguous().float(), real_B.contiguous().float())
                self._exit()
                return losses
            with self._enter():
                eng.stage_inputs(real_A, real_B)  # copy only: the step graph converts / pads the staged images itself
                eng.train_step()
                losses = eng.losses()
            self._exit()
            return losses
        return self._train_step_dp(eng, real_A, real_B)

    def _train_step_u8(self, real_A: torch.Tensor, real_B: torch.Tensor) -> Dict[str, float]:
        """uint8 interleaved RGB [N, H, W, 3] inputs (stand-in: train_step(from_uint8(a), from_uint8(b)))"""
        if self.sync.world_size != 1:
            raise NotImplementedError("uint8 inputs: single-process step only")
        n, h, w, c = real_A.shape
        probe = torch.empty(n, c, h, w, device="meta")
        eng = self._ensure_engine(probe)
        self._pool_step(eng)
        with self._enter():
            eng.stage_inputs_u8(real_A.contiguous(), real_B.contiguous())
            eng.train_step()
            losses = eng.losses()
        self._exit()
        return losses

    def _train_step_dp(self, eng, real_A, real_B) -> Dict[str, float]:
        self._train_step_dp_nosync(eng, real_A, real_B)
        with torch.cuda.stream(self.stream):
            losses = eng.losses()
        torch.cuda.current_stream().wait_stream(self.stream)
        return losses

    def _train_step_dp_nosync(self, eng, real_A, real_B) -> None:
        """one data-parallel step, fully asynchronous (no host read-back)"""
        main, comm = self.stream, self.comm_stream
        main.wait_stream(torch.cuda.current_stream())
        ev_G, ev_D, ev_done = torch.cuda.Event(), torch.cuda.Event(), torch.cuda.Event()
        if not self._dp_segmented:
            # merged schedule (CGB_SEG_STEP_NOOPT): forward, G-phase and D-phase backward as ONE graph with the D phase
            # in the shadow of the generator chains; then both all-reduces, each followed by its optimiser.
            # Measured on 2 x B200 at batch 1: faster than the segmented step, whose forwards / G phase / D phase
            # are separate graphs (the generator all-reduce overlapped the D phase there).
            with torch.cuda.stream(main):
                eng.stage_inputs(real_A, real_B)
                eng.run_segment(6)
                ev_G.record(main)
            with torch.cuda.stream(comm):
                comm.wait_event(ev_G)
                self.sync.all_reduce_(eng.grads[1])   # discriminators first: small, lets Adam(D) start early
                self.sync.all_reduce_(eng.grads[0])
                ev_D.record(comm)
                eng.run_segment(3)                    # Adam(G) * 1/world + bf16 weight refresh
                ev_done.record(comm)
            with torch.cuda.stream(main):
                main.wait_event(ev_D)
                eng.run_segment(4)                    # Adam(D) beside Adam(G)
            main.wait_event(ev_done)
            return
        with torch.cuda.stream(main):
            eng.stage_inputs(real_A, real_B)
            eng.run_segment(1)  # images, six forwards, G-phase backward (one CUDA graph)
            ev_G.record(main)
        with torch.cuda.stream(comm):
            comm.wait_event(ev_G)
            self.sync.all_reduce_(eng.grads[0])
            eng.run_segment(3)  # Adam(G) * 1/world + bf16 weight refresh
        with torch.cuda.stream(main):
            # needs only D weights and the pre-update fakes: overlaps the generator all-reduce + Adam
            eng.run_segment(2)
            ev_D.record(main)
        with torch.cuda.stream(comm):
            comm.wait_event(ev_D)
            self.sync.all_reduce_(eng.grads[1])
            eng.run_segment(4)
            ev_done.record(comm)
        main.wait_event(ev_done)

    # ---- introspection for tests ---------------------------------------------------------------------
    def grads(self, net_name: str) -> "OrderedDict[str, torch.Tensor]":
        net = dict(G_AB=0, G_BA=1, D_A=2, D_B=3)[net_name]
        return self.engine.grad_views(net)
