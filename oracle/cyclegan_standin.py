"""STAND-IN ORACLE (test infrastructure, NOT product code).

The nominal reference, EleutherAI/Unpaired-Image-Generation, contains no code
(`/root/reference/README.md:1` is a title line and the whole repository).  Per
BASELINE.json `north_star`, a minimal PyTorch-CPU implementation of the canonical
CycleGAN training step plays the role of the reference for correctness and CPU
timing.  This file IS that stand-in.  It follows the published CycleGAN design
(Zhu, Park, Isola, Efros, ICCV 2017): ResNet-9-block generators, 70x70 PatchGAN
discriminators, LSGAN + L1 cycle + L1 identity losses, Adam(2e-4, betas 0.5/0.999).

Parity pinning: the reference holds no golden vectors, so "parity unpinned" at the
reference; the only pins are this stand-in's own seeded outputs, frozen in
`tests/golden/` by `oracle/make_golden.py`.

Only `tests/`, `__graft_entry__.smoke()` and `bench.py`'s cpu_baseline / `--impl
reference` legs may import this module.  The product package
(`unpaired_image_generation_b200`) never does.

Two numeric modes:
  * fp32 (default)           -- the stand-in proper.
  * emulate_bf16=True        -- same graph with bf16 rounding inserted at exactly the
    points where the B200 pipeline stores a tensor in bf16 (see DESIGN.md "precision
    policy").  Used by the end-to-end parity tests, because randomly initialised
    InstanceNorm stacks amplify bf16 rounding beyond 1e-2 against pure fp32
    (SURVEY.md section 4.2).
"""
from __future__ import annotations

from collections import OrderedDict
from typing import Dict, Optional

import torch
import torch.nn as nn
import torch.nn.functional as F


# --------------------------------------------------------------------------------------
# bf16 storage-point emulation
# --------------------------------------------------------------------------------------
def _r(x: torch.Tensor) -> torch.Tensor:
    return x.to(torch.bfloat16).to(torch.float32)


class _RoundBoth(torch.autograd.Function):
    """value stored in bf16, and its gradient is stored in bf16 too."""

    @staticmethod
    def forward(ctx, x):
        return _r(x)

    @staticmethod
    def backward(ctx, g):
        return _r(g)


class _RoundFwd(torch.autograd.Function):
    """value stored in bf16; gradient consumed on the fly (never stored)."""

    @staticmethod
    def forward(ctx, x):
        return _r(x)

    @staticmethod
    def backward(ctx, g):
        return g


class _RoundGrad(torch.autograd.Function):
    """value untouched; its gradient is stored in bf16 (dgrad / act-backward outputs)."""

    @staticmethod
    def forward(ctx, x):
        return x.view_as(x)

    @staticmethod
    def backward(ctx, g):
        return _r(g)


class Precision:
    """q: both ways, qf: forward only, qg: gradient only. Identity in fp32 mode."""

    def __init__(self, emulate_bf16: bool = False):
        self.emulate_bf16 = emulate_bf16

    def q(self, x):
        return _RoundBoth.apply(x) if self.emulate_bf16 else x

    def qf(self, x):
        return _RoundFwd.apply(x) if self.emulate_bf16 else x

    def qg(self, x):
        return _RoundGrad.apply(x) if self.emulate_bf16 else x


_FP32 = Precision(False)


def _inorm(x: torch.Tensor) -> torch.Tensor:
    # InstanceNorm2d(affine=False, track_running_stats=False, eps=1e-5): biased variance
    return F.instance_norm(x, eps=1e-5)


def _init_conv(m: nn.Module) -> None:
    nn.init.normal_(m.weight, 0.0, 0.02)
    nn.init.zeros_(m.bias)


# --------------------------------------------------------------------------------------
# Generator: c7s1-64, d128, d256, 9 x R256, u128, u64, c7s1-3 + tanh
# --------------------------------------------------------------------------------------
class ResnetBlock(nn.Module):
    def __init__(self, ch: int):
        super().__init__()
        self.conv1 = nn.Conv2d(ch, ch, 3, bias=True)
        self.conv2 = nn.Conv2d(ch, ch, 3, bias=True)
        _init_conv(self.conv1)
        _init_conv(self.conv2)

    def forward(self, x: torch.Tensor, P: Precision = _FP32) -> torch.Tensor:
        h = P.qg(F.pad(x, (1, 1, 1, 1), mode="reflect"))
        h = P.q(F.conv2d(h, P.qf(self.conv1.weight), self.conv1.bias))
        h = P.qf(F.relu(_inorm(h)))
        h = P.qg(F.pad(h, (1, 1, 1, 1), mode="reflect"))
        h = P.q(F.conv2d(h, P.qf(self.conv2.weight), self.conv2.bias))
        return P.q(x + _inorm(h))


class Generator(nn.Module):
    """ResNet generator. forward(x:[N,3,H,W] in [-1,1]) -> [N,3,H,W] in (-1,1)."""

    def __init__(self, in_ch: int = 3, out_ch: int = 3, ngf: int = 64, n_blocks: int = 9):
        super().__init__()
        self.stem = nn.Conv2d(in_ch, ngf, 7, bias=True)
        self.down1 = nn.Conv2d(ngf, ngf * 2, 3, stride=2, padding=1, bias=True)
        self.down2 = nn.Conv2d(ngf * 2, ngf * 4, 3, stride=2, padding=1, bias=True)
        self.res = nn.ModuleList([ResnetBlock(ngf * 4) for _ in range(n_blocks)])
        self.up1 = nn.ConvTranspose2d(ngf * 4, ngf * 2, 3, stride=2, padding=1, output_padding=1, bias=True)
        self.up2 = nn.ConvTranspose2d(ngf * 2, ngf, 3, stride=2, padding=1, output_padding=1, bias=True)
        self.head = nn.Conv2d(ngf, out_ch, 7, bias=True)
        for m in (self.stem, self.down1, self.down2, self.up1, self.up2, self.head):
            _init_conv(m)

    def forward(self, x: torch.Tensor, P: Precision = _FP32, capture: Optional[dict] = None) -> torch.Tensor:
        def cap(name, t):
            if capture is not None:
                capture[name] = t
            return t

        h = P.qf(x)
        h = P.qg(F.pad(h, (3, 3, 3, 3), mode="reflect"))
        h = P.q(F.conv2d(h, P.qf(self.stem.weight), self.stem.bias))
        h = cap("stem", P.qf(F.relu(_inorm(h))))
        h = P.q(F.conv2d(P.qg(h), P.qf(self.down1.weight), self.down1.bias, stride=2, padding=1))
        h = cap("down1", P.qf(F.relu(_inorm(h))))
        h = P.q(F.conv2d(P.qg(h), P.qf(self.down2.weight), self.down2.bias, stride=2, padding=1))
        h = cap("down2", P.q(F.relu(_inorm(h))))
        for i, blk in enumerate(self.res):
            h = cap(f"res{i}", blk(h, P))
        h = P.q(F.conv_transpose2d(P.qg(h), P.qf(self.up1.weight), self.up1.bias, stride=2, padding=1, output_padding=1))
        h = cap("up1", P.qf(F.relu(_inorm(h))))
        h = P.q(F.conv_transpose2d(P.qg(h), P.qf(self.up2.weight), self.up2.bias, stride=2, padding=1, output_padding=1))
        h = cap("up2", P.qf(F.relu(_inorm(h))))
        h = P.qg(F.pad(h, (3, 3, 3, 3), mode="reflect"))
        h = P.qg(F.conv2d(h, P.qf(self.head.weight), self.head.bias))
        return cap("out", P.qf(torch.tanh(h)))


# --------------------------------------------------------------------------------------
# Discriminator: 70x70 PatchGAN  C64 - C128 - C256 - C512(s1) - C1(s1), 4x4 kernels
# --------------------------------------------------------------------------------------
class Discriminator(nn.Module):
    """forward(x:[N,3,H,W]) -> patch logits [N,1,H/8-2,W/8-2] (30x30 for 256x256)."""

    def __init__(self, in_ch: int = 3, ndf: int = 64, n_layers: int = 3):
        super().__init__()
        assert n_layers == 3, "the stand-in pins the canonical 70x70 PatchGAN (n_layers=3)"
        self.conv0 = nn.Conv2d(in_ch, ndf, 4, stride=2, padding=1, bias=True)
        self.conv1 = nn.Conv2d(ndf, ndf * 2, 4, stride=2, padding=1, bias=True)
        self.conv2 = nn.Conv2d(ndf * 2, ndf * 4, 4, stride=2, padding=1, bias=True)
        self.conv3 = nn.Conv2d(ndf * 4, ndf * 8, 4, stride=1, padding=1, bias=True)
        self.conv4 = nn.Conv2d(ndf * 8, 1, 4, stride=1, padding=1, bias=True)
        for m in (self.conv0, self.conv1, self.conv2, self.conv3, self.conv4):
            _init_conv(m)

    def forward(self, x: torch.Tensor, P: Precision = _FP32, capture: Optional[dict] = None) -> torch.Tensor:
        def cap(name, t):
            if capture is not None:
                capture[name] = t
            return t

        h = P.qg(P.qf(x))
        h = P.qg(F.conv2d(h, P.qf(self.conv0.weight), self.conv0.bias, stride=2, padding=1))
        h = cap("conv0", P.qf(F.leaky_relu(h, 0.2)))
        for name, conv, stride in (("conv1", self.conv1, 2), ("conv2", self.conv2, 2), ("conv3", self.conv3, 1)):
            h = P.q(F.conv2d(P.qg(h), P.qf(conv.weight), conv.bias, stride=stride, padding=1))
            h = cap(name, P.qf(F.leaky_relu(_inorm(h), 0.2)))
        h = P.q(F.conv2d(P.qg(h), P.qf(self.conv4.weight), self.conv4.bias, stride=1, padding=1))
        return cap("out", h)


# names of conv biases that feed a non-affine InstanceNorm: their gradient is
# mathematically zero (rounding noise in practice), see SURVEY.md section 4.2 item 3.
def dead_bias_names_generator(n_blocks: int = 9):
    names = ["stem.bias", "down1.bias", "down2.bias", "up1.bias", "up2.bias"]
    for i in range(n_blocks):
        names += [f"res.{i}.conv1.bias", f"res.{i}.conv2.bias"]
    return names


def dead_bias_names_discriminator():
    return ["conv1.bias", "conv2.bias", "conv3.bias"]


# --------------------------------------------------------------------------------------
# Input pipeline and learning-rate schedule of the canonical recipe (SURVEY.md section 8 f, items 2 and 3)
# --------------------------------------------------------------------------------------
def from_uint8(img_u8_hwc: torch.Tensor) -> torch.Tensor:
    """uint8 interleaved RGB [N, H, W, 3] -> float32 [N, 3, H, W] in [-1, 1]: ToTensor + Normalize(0.5, 0.5)."""
    assert img_u8_hwc.dtype == torch.uint8 and img_u8_hwc.dim() == 4 and img_u8_hwc.shape[-1] == 3
    return (img_u8_hwc.permute(0, 3, 1, 2).to(torch.float32) * (1.0 / 127.5) - 1.0).contiguous()


def to_uint8(img: torch.Tensor) -> torch.Tensor:
    """float [N, 3, H, W] in [-1, 1] -> uint8 interleaved RGB [N, H, W, 3]: clamp(round((x + 1) * 127.5), 0, 255)."""
    return ((img.to(torch.float32) + 1.0) * 127.5).round().clamp(0, 255).to(torch.uint8).permute(0, 2, 3, 1).contiguous()


def linear_decay_lr(base_lr: float, epoch: int, n_epochs: int = 100, n_epochs_decay: int = 100) -> float:
    """Constant for `n_epochs`, then linear decay to zero over `n_epochs_decay` (the canonical 'linear' policy):
    lr = base_lr * (1 - max(0, epoch + 1 - n_epochs) / (n_epochs_decay + 1)), epoch counted from 0."""
    return base_lr * (1.0 - max(0, epoch + 1 - n_epochs) / float(n_epochs_decay + 1))


class PoolDecisions:
    """The random decisions of the canonical ImagePool.query, one per image: (store, ret).
    ret >= 0: the discriminator sees the OLD pool[ret] instead of the new fake; store >= 0: the new fake is written
    to pool[store] afterwards.  Filling phase: (count, -1); then with probability 0.5 swap a random slot (i, i),
    else pass the new image through (-1, -1)."""

    def __init__(self, pool_size: int, seed: int = 0):
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


class ImagePool:
    """History of generated images shown to a discriminator (pool_size 50 in the canonical recipe)."""

    def __init__(self, pool_size: int, seed: int = 0):
        self.decisions = PoolDecisions(pool_size, seed)
        self.images = [None] * max(0, int(pool_size))

    def query(self, images: torch.Tensor) -> torch.Tensor:
        out = []
        for img in images.detach():
            store, ret = self.decisions.next()
            cur = img.clone()
            out.append(self.images[ret].clone() if ret >= 0 else cur)
            if store >= 0:
                self.images[store] = cur
        return torch.stack(out)


# --------------------------------------------------------------------------------------
# Training step
# --------------------------------------------------------------------------------------
class CycleGANTrainer:
    """Canonical CycleGAN optimisation step (constant lr unless set_lr is driven; image history pool when
    pool_size > 0: the discriminators then see pool.query(fake), pools seeded pool_seed / pool_seed + 1).

    train_step(real_A, real_B) -> dict of python floats:
      loss_G, loss_G_A, loss_G_B, loss_cycle_A, loss_cycle_B, loss_idt_A, loss_idt_B,
      loss_D_A, loss_D_B
    D_A judges domain-B images (fake_B = G_AB(real_A)); D_B judges domain-A images.
    """

    LOSS_KEYS = ("loss_G", "loss_G_A", "loss_G_B", "loss_cycle_A", "loss_cycle_B",
                 "loss_idt_A", "loss_idt_B", "loss_D_A", "loss_D_B")

    def __init__(self, G_AB: Generator, G_BA: Generator, D_A: Discriminator, D_B: Discriminator,
                 lambda_A: float = 10.0, lambda_B: float = 10.0, lambda_idt: float = 0.5,
                 lr: float = 2e-4, betas=(0.5, 0.999), eps: float = 1e-8,
                 emulate_bf16: bool = False, pool_size: int = 0, pool_seed: int = 0):
        self.G_AB, self.G_BA, self.D_A, self.D_B = G_AB, G_BA, D_A, D_B
        self.pool_B = ImagePool(pool_size, pool_seed) if pool_size > 0 else None      # fake_B history, for D_A
        self.pool_A = ImagePool(pool_size, pool_seed + 1) if pool_size > 0 else None  # fake_A history, for D_B
        self.lambda_A, self.lambda_B, self.lambda_idt = lambda_A, lambda_B, lambda_idt
        self.P = Precision(emulate_bf16)
        self.opt_G = torch.optim.Adam(list(G_AB.parameters()) + list(G_BA.parameters()), lr=lr, betas=betas, eps=eps)
        self.opt_D = torch.optim.Adam(list(D_A.parameters()) + list(D_B.parameters()), lr=lr, betas=betas, eps=eps)
        self.last_images: Dict[str, torch.Tensor] = {}

    def set_lr(self, lr: float) -> None:
        """learning rate of both optimisers from the next step on (drive it with linear_decay_lr)"""
        for opt in (self.opt_G, self.opt_D):
            for group in opt.param_groups:
                group["lr"] = float(lr)

    # -- the six generated images, for activation parity -------------------------------
    @torch.no_grad()
    def forward_only(self, real_A: torch.Tensor, real_B: torch.Tensor) -> "OrderedDict[str, torch.Tensor]":
        P = self.P
        out = OrderedDict()
        out["fake_B"] = self.G_AB(real_A, P)
        out["rec_A"] = self.G_BA(out["fake_B"], P)
        out["fake_A"] = self.G_BA(real_B, P)
        out["rec_B"] = self.G_AB(out["fake_A"], P)
        out["idt_A"] = self.G_AB(real_B, P)
        out["idt_B"] = self.G_BA(real_A, P)
        return out

    @staticmethod
    def _mse_to(pred: torch.Tensor, target: float) -> torch.Tensor:
        return ((pred - target) ** 2).mean()

    def compute_G_losses(self, real_A, real_B):
        P = self.P
        fake_B = self.G_AB(real_A, P)
        rec_A = self.G_BA(fake_B, P)
        fake_A = self.G_BA(real_B, P)
        rec_B = self.G_AB(fake_A, P)
        idt_A = self.G_AB(real_B, P)
        idt_B = self.G_BA(real_A, P)
        tA, tB = P.qf(real_A), P.qf(real_B)  # targets are the bf16-stored images on B200
        L = OrderedDict()
        L["loss_idt_A"] = F.l1_loss(idt_A, tB) * self.lambda_B * self.lambda_idt
        L["loss_idt_B"] = F.l1_loss(idt_B, tA) * self.lambda_A * self.lambda_idt
        L["loss_G_A"] = self._mse_to(self.D_A(fake_B, P), 1.0)
        L["loss_G_B"] = self._mse_to(self.D_B(fake_A, P), 1.0)
        L["loss_cycle_A"] = F.l1_loss(rec_A, tA) * self.lambda_A
        L["loss_cycle_B"] = F.l1_loss(rec_B, tB) * self.lambda_B
        L["loss_G"] = (L["loss_G_A"] + L["loss_G_B"] + L["loss_cycle_A"] + L["loss_cycle_B"]
                       + L["loss_idt_A"] + L["loss_idt_B"])
        imgs = dict(fake_B=fake_B, rec_A=rec_A, fake_A=fake_A, rec_B=rec_B, idt_A=idt_A, idt_B=idt_B)
        return L, imgs

    def compute_D_loss(self, D: Discriminator, real: torch.Tensor, fake: torch.Tensor) -> torch.Tensor:
        P = self.P
        return 0.5 * (self._mse_to(D(real, P), 1.0) + self._mse_to(D(fake.detach(), P), 0.0))

    def backward_only(self, real_A, real_B) -> Dict[str, float]:
        """Forward + both backward passes, no optimizer step (gradients left in .grad)."""
        for p in list(self.D_A.parameters()) + list(self.D_B.parameters()):
            p.requires_grad_(False)
        self.opt_G.zero_grad(set_to_none=True)
        L, imgs = self.compute_G_losses(real_A, real_B)
        L["loss_G"].backward()
        for p in list(self.D_A.parameters()) + list(self.D_B.parameters()):
            p.requires_grad_(True)
        self.opt_D.zero_grad(set_to_none=True)
        fake_B = self.pool_B.query(imgs["fake_B"]) if self.pool_B else imgs["fake_B"]
        fake_A = self.pool_A.query(imgs["fake_A"]) if self.pool_A else imgs["fake_A"]
        loss_D_A = self.compute_D_loss(self.D_A, real_B, fake_B)
        loss_D_A.backward()
        loss_D_B = self.compute_D_loss(self.D_B, real_A, fake_A)
        loss_D_B.backward()
        self.last_images = {k: v.detach() for k, v in imgs.items()}
        out = {k: float(v.detach()) for k, v in L.items()}
        out["loss_D_A"] = float(loss_D_A.detach())
        out["loss_D_B"] = float(loss_D_B.detach())
        return {k: out[k] for k in self.LOSS_KEYS}

    def train_step(self, real_A: torch.Tensor, real_B: torch.Tensor) -> Dict[str, float]:
        # G phase: D frozen, gradients flow through D into G
        for p in list(self.D_A.parameters()) + list(self.D_B.parameters()):
            p.requires_grad_(False)
        self.opt_G.zero_grad(set_to_none=True)
        L, imgs = self.compute_G_losses(real_A, real_B)
        L["loss_G"].backward()
        self.opt_G.step()
        # D phase: fakes computed BEFORE the G update, detached
        for p in list(self.D_A.parameters()) + list(self.D_B.parameters()):
            p.requires_grad_(True)
        self.opt_D.zero_grad(set_to_none=True)
        fake_B = self.pool_B.query(imgs["fake_B"]) if self.pool_B else imgs["fake_B"]
        fake_A = self.pool_A.query(imgs["fake_A"]) if self.pool_A else imgs["fake_A"]
        loss_D_A = self.compute_D_loss(self.D_A, real_B, fake_B)
        loss_D_A.backward()
        loss_D_B = self.compute_D_loss(self.D_B, real_A, fake_A)
        loss_D_B.backward()
        self.opt_D.step()
        self.last_images = {k: v.detach() for k, v in imgs.items()}
        out = {k: float(v.detach()) for k, v in L.items()}
        out["loss_D_A"] = float(loss_D_A.detach())
        out["loss_D_B"] = float(loss_D_B.detach())
        return {k: out[k] for k in self.LOSS_KEYS}


# --------------------------------------------------------------------------------------
# Deterministic construction + synthetic data
# --------------------------------------------------------------------------------------
def build_models(seed: int = 0, n_blocks: int = 9):
    torch.manual_seed(seed)
    G_AB = Generator(n_blocks=n_blocks)
    G_BA = Generator(n_blocks=n_blocks)
    D_A = Discriminator()
    D_B = Discriminator()
    return G_AB, G_BA, D_A, D_B


def synthetic_pair(batch: int, size: int, seed: int = 1234):
    g = torch.Generator().manual_seed(seed)
    real_A = torch.rand(batch, 3, size, size, generator=g) * 2 - 1
    real_B = torch.rand(batch, 3, size, size, generator=g) * 2 - 1
    return real_A, real_B
