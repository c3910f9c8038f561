"""CycleGANTrainer with the stand-in's API (oracle/cyclegan_standin.py:278): same constructor
arguments, `train_step(real_A, real_B) -> dict`, `forward_only`, `backward_only`, `set_lr`.  One process per
GPU.  With torch.distributed initialised the step is data parallel: the merged step graph announces gradient
buckets as they become final (discriminators first, then each generator from the head towards the stem) and the
all-reduce of a bucket, followed by Adam on that range, runs on a communication stream while the rest of the backward
pass is still executing; the bf16 weight refresh follows the step.  Initial parameters and optimiser state are
broadcast from rank 0."""
from __future__ import annotations

import os

from collections import OrderedDict
from typing import Dict, List, Optional, Tuple

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
    _DEC_RING = 4  # pinned pool-decision buffers in flight (each guarded by an event)

    def __init__(self, G_AB: Generator, G_BA: Generator, D_A: Discriminator, D_B: Discriminator,
                 lambda_A: float = 10.0, lambda_B: float = 10.0, lambda_idt: float = 0.5, lr: float = 2e-4,
                 betas=(0.5, 0.999), eps: float = 1e-8, process_group=None, pool_size: int = 0, pool_seed: int = 0,
                 precision: str = "bf16"):
        """precision: "bf16" (the product path) or "fp32" (deterministic validation mode: same schedule, fp32
        activations, fp64 accumulation; matches the fp32 stand-in to 1e-5 on activations and losses)"""
        self.G_AB, self.G_BA, self.D_A, self.D_B = G_AB, G_BA, D_A, D_B
        if G_AB.n_blocks != G_BA.n_blocks:
            raise ValueError("both generators must have the same number of residual blocks")
        self._hyper = dict(lambda_A=lambda_A, lambda_B=lambda_B, lambda_idt=lambda_idt, lr=lr, betas=betas, eps=eps)
        self.precision = precision
        self.sync = GradSync(process_group)
        self.pool_size = int(pool_size)
        self._pools = (_PoolDecisions(pool_size, pool_seed), _PoolDecisions(pool_size, pool_seed + 1))  # fake_B, fake_A
        self.engine: Optional[_engine.StepEngine] = None
        self.stream: Optional[torch.cuda.Stream] = None
        self.comm_stream: Optional[torch.cuda.Stream] = None
        self._dp_segmented = bool(os.environ.get("CGB_DP_SEGMENTED"))  # the round-1 segmented data-parallel step
        self._dp_overlap = os.environ.get("CGB_DP_OVERLAP", "1") != "0"
        self._buckets: List[dict] = []  # engine.grad_bucket_plan(): gradient ranges in the order they become final
        self._dec_ring: List[Tuple[torch.Tensor, Optional[torch.cuda.Event]]] = []
        self._dec_next = 0

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
        eng = _engine.StepEngine(batch, h, self.G_AB.n_blocks, pool_size=self.pool_size, precision=self.precision,
                                 **self._hyper)
        for net, mod in enumerate((self.G_AB, self.G_BA, self.D_A, self.D_B)):
            mod._attach(eng, net)
        if self.sync.world_size > 1:
            # every rank must start from the same weights and optimiser state: rank 0's (modules are seeded per
            # process and load_state_dict is per rank, so nothing else guarantees it)
            for t in eng.params + eng.exp_avg + eng.exp_avg_sq:
                self.sync.broadcast_(t)
            steps = torch.tensor([eng.step_count(0), eng.step_count(1)], dtype=torch.int64, device=eng.device)
            self.sync.broadcast_(steps)
            for g in range(2):
                eng.set_step_count(g, int(steps[g]))
        eng.refresh_weights(0)
        eng.refresh_weights(1)
        eng.set_grad_scale(self.sync.grad_scale)
        self.engine = eng
        self.stream = torch.cuda.Stream(device=eng.device)
        # the communication stream runs at high priority: a collective only progresses once EVERY rank's NCCL CTAs are
        # resident, so its blocks should not queue behind the step's pending conv CTAs (CGB_DP_COMM_PRIO=0: default priority)
        prio = int(os.environ.get("CGB_DP_COMM_PRIO", "-1"))
        self.comm_stream = torch.cuda.Stream(device=eng.device, priority=prio)
        self._buckets = eng.grad_bucket_plan()
        if self.pool_size > 0:
            self._dec_ring = [(torch.empty(2, batch, 2, dtype=torch.int32).pin_memory(), None)
                              for _ in range(self._DEC_RING)]
        if getattr(self, "_lr", None) is not None:  # set_lr() before the first step
            eng.set_lr(self._lr)
        return eng

    def _check_pair(self, eng, real_A: torch.Tensor, real_B: torch.Tensor, dtype, shape) -> None:
        for name, t in (("real_A", real_A), ("real_B", real_B)):
            if t.dtype != dtype or tuple(t.shape) != shape:
                raise ValueError(f"{name}: expected {dtype} {shape}, got {t.dtype} {tuple(t.shape)}")
        if real_A.device != real_B.device:
            raise ValueError(f"real_A is on {real_A.device} but real_B is on {real_B.device}")

    def _pool_step(self, eng) -> None:
        """this step's image-pool decisions (one per image and pool), sent ahead of the D phase.  The pinned staging
        buffers form a small ring, each guarded by an event recorded after its copy was enqueued: a buffer is only
        rewritten once the copy that read it has run (the raw cudaMemcpyAsync is invisible to torch's allocator)."""
        if self.pool_size <= 0:
            return
        slot = self._dec_next
        dec, ev = self._dec_ring[slot]
        if ev is not None:
            ev.synchronize()
        for side in range(2):
            for n in range(eng.batch):
                store, ret = self._pools[side].next()
                dec[side, n, 0], dec[side, n, 1] = store, ret
        with torch.cuda.stream(self.stream):
            eng.set_pool_decisions(dec)
            ev = torch.cuda.Event()
            ev.record(self.stream)
        self._dec_ring[slot] = (dec, ev)
        self._dec_next = (slot + 1) % self._DEC_RING

    def _enter(self):
        self.stream.wait_stream(torch.cuda.current_stream())
        return torch.cuda.stream(self.stream)

    def _exit(self):
        torch.cuda.current_stream().wait_stream(self.stream)

    def _stepped(self):
        for m in (self.G_AB, self.G_BA, self.D_A, self.D_B):
            m._trained_steps += 1

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
        """one optimisation step; inputs: fp32 [N, 3, H, H] in [-1, 1] (device, or host -- pinned for async copies)
        or uint8 interleaved RGB [N, H, H, 3] (stand-in: train_step(from_uint8(a), from_uint8(b)))"""
        u8 = real_A.dtype == torch.uint8
        if u8:
            n, h, w, c = real_A.shape
            eng = self._ensure_engine(torch.empty(n, c, h, w, device="meta"))
            self._check_pair(eng, real_A, real_B, torch.uint8, (eng.batch, eng.size, eng.size, 3))
            stage = lambda: eng.stage_inputs_u8(real_A.contiguous(), real_B.contiguous())
        else:
            eng = self._ensure_engine(real_A)
            shape = (eng.batch, 3, eng.size, eng.size)
            if tuple(real_A.shape) != shape or tuple(real_B.shape) != shape:
                raise ValueError(f"expected two inputs of shape {shape}, got {tuple(real_A.shape)} and {tuple(real_B.shape)}")
            if real_A.device != real_B.device:
                raise ValueError(f"real_A is on {real_A.device} but real_B is on {real_B.device}")
            stage = lambda: eng.stage_inputs(real_A, real_B)
        self._pool_step(eng)
        if self.sync.world_size > 1:
            self._train_step_dp_nosync(eng, stage)
            with torch.cuda.stream(self.stream):
                losses = eng.losses()
            torch.cuda.current_stream().wait_stream(self.stream)
        elif not u8 and real_A.device.type == "cpu":
            # both tensors were checked above: fp32 host buffers go straight to the C ABI (H2D + step + D2H)
            with self._enter():
                losses = eng.train_step_host(real_A.contiguous().float(), real_B.contiguous().float())
            self._exit()
        else:
            with self._enter():
                stage()               # copy only: the step graph converts / pads the staged images itself
                eng.train_step()
                losses = eng.losses()
            self._exit()
        self._stepped()
        return losses

    def _train_step_dp_nosync(self, eng, stage) -> None:
        """one data-parallel step, fully asynchronous (no host read-back)"""
        main, comm = self.stream, self.comm_stream
        main.wait_stream(torch.cuda.current_stream())
        ev_G, ev_D, ev_done = torch.cuda.Event(), torch.cuda.Event(), torch.cuda.Event()
        if not self._dp_segmented:
            # merged schedule (CGB_SEG_STEP_NOOPT): forward, G-phase and D-phase backward as ONE graph with the D phase
            # in the shadow of the generator chains.  The graph records an external event per gradient bucket; the
            # communication stream waits on each in turn and all-reduces that range while the graph keeps running.
            with torch.cuda.stream(main):
                stage()
                eng.run_segment(6)
                ev_G.record(main)
            with torch.cuda.stream(comm):
                if self._dp_overlap:
                    # bucket by bucket: wait until the range is final, all-reduce it, Adam on it -- all while the step
                    # graph keeps running (nothing left in the step reads the fp32 masters of a finished bucket; the
                    # bf16 packs the kernels do read are refreshed after the step)
                    # The bf16 packs of a generator bucket are refreshed once the NEXT bucket of the same generator is
                    # final (the chain has then left those layers; the discriminators' right away: nothing reads them
                    # after their phase); the last bucket of each generator after the whole step.
                    first = [True, True]
                    pending = {}  # net -> bucket whose packs wait for the next bucket of that network
                    for b in self._buckets:
                        eng.wait_grad_bucket(b["index"])
                        prev = pending.pop(b["net"], None)
                        if prev is not None:
                            eng.refresh_weights_layers(prev["net"], prev["layer_lo"], prev["layer_hi"])
                        g, off, numel = b["group"], b["offset"], b["numel"]
                        self.sync.all_reduce_(eng.grads[g][off:off + numel])
                        eng.adam_range(g, off, numel, first[g])
                        first[g] = False
                        if b["net"] >= 0:
                            pending[b["net"]] = b
                        elif g == 1:
                            eng.refresh_weights(1)
                    comm.wait_event(ev_G)             # the whole step: the conv kernels are done with the old bf16 weights
                    for prev in pending.values():
                        eng.refresh_weights_layers(prev["net"], prev["layer_lo"], prev["layer_hi"])
                    if any(b["net"] < 0 and b["group"] == 0 for b in self._buckets):
                        eng.refresh_weights(0)        # (single-bucket / validation mode: the whole generator group)
                    ev_done.record(comm)
                else:
                    comm.wait_event(ev_G)
                    self.sync.all_reduce_(eng.grads[1])   # discriminators first: small, lets Adam(D) start early
                    self.sync.all_reduce_(eng.grads[0])
                    ev_D.record(comm)
                    eng.run_segment(3)                # Adam(G) * 1/world + bf16 weight refresh
                    ev_done.record(comm)
            if not self._dp_overlap:
                with torch.cuda.stream(main):
                    main.wait_event(ev_D)
                    eng.run_segment(4)                # Adam(D) beside Adam(G)
            main.wait_event(ev_done)
            return
        with torch.cuda.stream(main):
            stage()
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

    # ---- checkpoint / resume ----------------------------------------------------------------------------
    def state_dict(self) -> Dict[str, object]:
        """everything a resumed run needs: the four modules' weights and both Adam states (moments + step counters).
        Stand-in: {net.state_dict()} + opt_G.state_dict() + opt_D.state_dict() (oracle/cyclegan_standin.py:300)."""
        if self.engine is None:
            raise RuntimeError("state_dict() needs a bound engine: run a step (or forward_only) first")
        torch.cuda.synchronize(self.engine.device)
        return {"G_AB": self.G_AB.state_dict(), "G_BA": self.G_BA.state_dict(), "D_A": self.D_A.state_dict(),
                "D_B": self.D_B.state_dict(), "optimizer": self.engine.optimizer_state(),
                "lr": getattr(self, "_lr", None),
                "pools": [(p.count, p.rng.getstate()) for p in self._pools]}

    def load_state_dict(self, state: Dict[str, object], like: Optional[torch.Tensor] = None) -> None:
        """restore `state_dict()`.  The engine is created for inputs shaped like `like` ([N, 3, H, H]) if it does not
        exist yet.  The image history pool's IMAGES are not part of the state (as in the canonical recipe)."""
        if self.engine is None:
            if like is None:
                raise RuntimeError("load_state_dict() before the first step needs `like` (an input-shaped tensor)")
            self._ensure_engine(like)
        for name in ("G_AB", "G_BA", "D_A", "D_B"):
            getattr(self, name).load_state_dict(state[name])
        torch.cuda.synchronize(self.engine.device)
        self.engine.load_optimizer_state(state["optimizer"])
        if state.get("lr") is not None:
            self.set_lr(state["lr"])
        for p, (count, rng) in zip(self._pools, state.get("pools", [])):
            p.count = count
            p.rng.setstate(rng)
        torch.cuda.synchronize(self.engine.device)

    # ---- introspection for tests ---------------------------------------------------------------------
    def grads(self, net_name: str) -> "OrderedDict[str, torch.Tensor]":
        net = dict(G_AB=0, G_BA=1, D_A=2, D_B=3)[net_name]
        return self.engine.grad_views(net)
