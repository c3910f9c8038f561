"""CPU tests of the STAND-IN oracle against its frozen golden vectors.

The reference ships no tests or fixtures (/root/reference/README.md:1 is the whole
repo); these goldens were produced by oracle/make_golden.py and pin the stand-in.
"""
import math

import numpy as np
import pytest
import torch

from oracle.cyclegan_standin import (CycleGANTrainer, Discriminator, Generator, Precision,
                                     build_models, dead_bias_names_discriminator,
                                     dead_bias_names_generator, synthetic_pair)


def _close(a, b, rel=2e-4, abs_=1e-5):
    return abs(a - b) <= abs_ + rel * abs(b)


def test_shapes_and_param_counts(golden):
    G = Generator()
    D = Discriminator()
    assert sum(p.numel() for p in G.parameters()) == 11_378_179 == golden["param_counts"]["G"]
    assert sum(p.numel() for p in D.parameters()) == 2_764_737 == golden["param_counts"]["D"]
    assert len(list(G.parameters())) == 48 and len(list(D.parameters())) == 10
    x = torch.zeros(2, 3, 64, 64)
    with torch.no_grad():
        assert G(x).shape == (2, 3, 64, 64)
        assert D(x).shape == (2, 1, 6, 6)
        assert D(torch.zeros(1, 3, 256, 256)).shape == (1, 1, 30, 30)
    assert len(dead_bias_names_generator()) == 23 and len(dead_bias_names_discriminator()) == 3


def test_init_is_seeded_and_normal():
    a = build_models(seed=0)[0]
    b = build_models(seed=0)[0]
    for (n1, p1), (n2, p2) in zip(a.named_parameters(), b.named_parameters()):
        assert n1 == n2 and torch.equal(p1, p2)
    w = a.res[3].conv2.weight
    assert abs(float(w.std()) - 0.02) < 5e-4 and float(a.res[3].conv2.bias.abs().max()) == 0.0


@pytest.mark.parametrize("case", ["fp32_64", "bf16emu_64", "fp32_64_b2"])
def test_golden_small(case, golden, golden_samples):
    g = golden["cases"][case]
    G_AB, G_BA, D_A, D_B = build_models(seed=0)
    tr = CycleGANTrainer(G_AB, G_BA, D_A, D_B, emulate_bf16=g["emulate_bf16"])
    real_A, real_B = synthetic_pair(g["batch"], g["size"], seed=1234)
    imgs = tr.forward_only(real_A, real_B)
    for k, st in g["images"].items():
        assert _close(float(imgs[k].double().abs().sum()), st["abs_sum"], rel=1e-3), k
        key = f"{case}.img_{k}"
        if key in golden_samples:
            np.testing.assert_allclose(imgs[k][0, :, ::8, ::8].numpy(), golden_samples[key],
                                       rtol=0, atol=2e-2 if g["emulate_bf16"] else 2e-4)
    losses = tr.backward_only(real_A, real_B)
    for k, v in g["losses_step0"].items():
        assert _close(losses[k], v, rel=2e-3 if g["emulate_bf16"] else 2e-4), (k, losses[k], v)
    if not g["emulate_bf16"]:
        for name, st in g["grads_step0"].items():
            net, pname = name.split(".", 1)
            p = dict(dict(G_AB=G_AB, G_BA=G_BA, D_A=D_A, D_B=D_B)[net].named_parameters())[pname]
            assert _close(float(p.grad.double().norm()), st["l2"], rel=2e-3), name
    for step_golden in g["losses_steps"]:
        got = tr.train_step(real_A, real_B)
        for k, v in step_golden.items():
            assert _close(got[k], v, rel=5e-3 if g["emulate_bf16"] else 1e-3), (k, got[k], v)


def test_golden_256_losses(golden):
    g = golden["cases"]["fp32_256"]
    G_AB, G_BA, D_A, D_B = build_models(seed=0)
    tr = CycleGANTrainer(G_AB, G_BA, D_A, D_B)
    real_A, real_B = synthetic_pair(1, 256, seed=1234)
    losses = tr.backward_only(real_A, real_B)
    for k, v in g["losses_step0"].items():
        assert _close(losses[k], v, rel=2e-4), (k, losses[k], v)


def test_adam_first_step_is_sign_of_gradient():
    # |dw| == lr on the first Adam step for live weights
    G_AB, G_BA, D_A, D_B = build_models(seed=0)
    tr = CycleGANTrainer(G_AB, G_BA, D_A, D_B)
    w0 = D_A.conv4.weight.detach().clone()
    real_A, real_B = synthetic_pair(1, 32, seed=1)
    tr.train_step(real_A, real_B)
    dw = (D_A.conv4.weight.detach() - w0).abs() / 2e-4
    assert 0.9 < float(dw.median()) < 1.01


def test_dead_biases_have_noise_gradients():
    G_AB, G_BA, D_A, D_B = build_models(seed=0)
    tr = CycleGANTrainer(G_AB, G_BA, D_A, D_B)
    real_A, real_B = synthetic_pair(1, 32, seed=1)
    tr.backward_only(real_A, real_B)
    params = dict(G_AB.named_parameters())
    for n in dead_bias_names_generator():
        assert float(params[n].grad.abs().max()) < 1e-4
    assert float(params["head.bias"].grad.abs().max()) > 1e-3


def test_bf16_emulation_rounds_storage_points():
    P = Precision(True)
    x = torch.randn(1000, requires_grad=True)
    y = P.q(x)
    assert torch.equal(y.detach(), x.detach().to(torch.bfloat16).float())
    (y * torch.randn(1000)).sum().backward()
    assert torch.equal(x.grad, x.grad.to(torch.bfloat16).float())
    # forward-only / grad-only variants
    x2 = torch.randn(100, requires_grad=True)
    g = torch.randn(100)
    (P.qf(x2) * g).sum().backward()
    assert torch.equal(x2.grad, g)
    x3 = torch.randn(100, requires_grad=True)
    y3 = P.qg(x3)
    assert torch.equal(y3.detach(), x3.detach())
    (y3 * g).sum().backward()
    assert torch.equal(x3.grad, g.to(torch.bfloat16).float())


def test_dp_equivalence_batch2_equals_two_ranks():
    """InstanceNorm is per-sample, so batch-2 gradients == mean of two batch-1 gradients."""
    real_A, real_B = synthetic_pair(2, 32, seed=7)
    G_AB, G_BA, D_A, D_B = build_models(seed=0)
    tr = CycleGANTrainer(G_AB, G_BA, D_A, D_B)
    tr.backward_only(real_A, real_B)
    full = G_AB.res[2].conv1.weight.grad.clone()
    acc = torch.zeros_like(full)
    for i in range(2):
        G2 = build_models(seed=0)
        tr2 = CycleGANTrainer(*G2)
        tr2.backward_only(real_A[i:i + 1], real_B[i:i + 1])
        acc += G2[0].res[2].conv1.weight.grad
    rel = float((acc / 2 - full).norm() / full.norm())
    assert rel < 5e-3, rel  # fp32 summation-order noise through sign() of the L1 losses


def test_from_uint8_and_linear_decay_schedule():
    """input pipeline and LR policy of the canonical recipe (SURVEY.md section 8 f), as restated by the stand-in"""
    from oracle.cyclegan_standin import from_uint8, linear_decay_lr
    u8 = torch.tensor([[[[0, 127, 255], [128, 64, 1]]]], dtype=torch.uint8)  # [1, 1, 2, 3]
    x = from_uint8(u8)
    assert x.shape == (1, 3, 1, 2) and x.dtype == torch.float32
    assert float(x[0, 0, 0, 0]) == -1.0 and float(x[0, 2, 0, 0]) == 1.0
    assert abs(float(x[0, 1, 0, 0]) - (127 / 127.5 - 1)) < 1e-7
    assert linear_decay_lr(2e-4, 0) == 2e-4 and linear_decay_lr(2e-4, 99) == 2e-4
    assert abs(linear_decay_lr(2e-4, 100) - 2e-4 * (1 - 1 / 101)) < 1e-12
    assert abs(linear_decay_lr(2e-4, 199) - 2e-4 * (1 - 100 / 101)) < 1e-12


def test_set_lr_changes_the_next_adam_step():
    from oracle.cyclegan_standin import CycleGANTrainer, build_models, synthetic_pair
    nets = build_models(seed=0, n_blocks=1) if "n_blocks" in build_models.__code__.co_varnames else build_models(seed=0)
    tr = CycleGANTrainer(*nets)
    a, b = synthetic_pair(1, 32, seed=3)
    tr.set_lr(0.0)
    w0 = nets[0].head.weight.detach().clone()
    tr.train_step(a, b)
    assert torch.equal(nets[0].head.weight.detach(), w0)  # lr 0: no update
    tr.set_lr(1e-3)
    tr.train_step(a, b)
    step = (nets[0].head.weight.detach() - w0).abs().max()
    assert 0 < float(step) <= 1.01e-3


def test_image_pool_semantics_and_decisions_match_the_product():
    """canonical ImagePool: fill, then swap a random slot with probability 0.5; the product's host-side decision
    generator (trainer._PoolDecisions) restates the stand-in's PoolDecisions and must produce the same sequence"""
    from oracle.cyclegan_standin import ImagePool, PoolDecisions
    from unpaired_image_generation_b200.trainer import _PoolDecisions
    a, b = PoolDecisions(3, seed=5), _PoolDecisions(3, 5)
    seq = [a.next() for _ in range(200)]
    assert seq == [b.next() for _ in range(200)]
    assert seq[:3] == [(0, -1), (1, -1), (2, -1)]
    swaps = [s for s in seq[3:] if s != (-1, -1)]
    assert 60 < len(swaps) < 140 and all(s[0] == s[1] and 0 <= s[0] < 3 for s in swaps)
    pool = ImagePool(2, seed=1)
    imgs = [torch.full((1, 3, 2, 2), float(i)) for i in range(12)]
    outs = [float(pool.query(x)[0, 0, 0, 0]) for x in imgs]
    assert outs[:2] == [0.0, 1.0]                       # filling: images pass through
    assert all(o <= i for i, o in enumerate(outs))      # never an image from the future
    assert any(o < i for i, o in enumerate(outs))       # history is used
    assert PoolDecisions(0).next() == (-1, -1)
