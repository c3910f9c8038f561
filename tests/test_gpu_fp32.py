"""GPU tests of the fp32 VALIDATION MODE (north_star: "1e-5 for an fp32 validation mode"); run on a B200 via
`pytest -m gpu`.  Same engine, programs and schedule as the bf16 product path, fp32 activations, fp64 accumulation,
deterministic kernels (csrc/fp32_path.h).  Checker: the fp32 stand-in (oracle/cyclegan_standin.py), run LIVE on the
host cores at 256x256 batch 1 (BASELINE.json configs[1]), and the committed float64 golden vectors
(tests/golden/standin_fp64_256.npz, oracle/make_golden_fp64.py).

Tolerances, with the measured noise floor of the fp32 stand-in against its own float64 run
(profiles/r02_fp32_noise_floor.txt) in brackets:
  * single layers vs a float64 torch reference:          <= 2e-6 relative L2
  * the nine losses vs the fp32 stand-in:                <= 1e-5   [1e-7]
  * the six images vs the fp32 stand-in:                 <= 1e-5   [1e-6 .. 7e-6 on the rec images]
  * the six images vs the float64 golden samples:        <= 1e-5
  * weight gradients vs the fp32 stand-in / float64:     <= 1e-2   [median 1.8e-3, worst 4e-3: ReLU and L1-sign
    decisions that flip under 1e-7 perturbations move a gradient by O(sqrt(fraction flipped)); no implementation can
    match the fp32 stand-in's gradients more tightly than the stand-in matches exact arithmetic]
  * weights after one Adam step vs the fp32 stand-in:    <= 2e-3   [first Adam step = lr * sign(g): elements whose
    tiny gradient changes sign move by 2 lr = 2 % of a weight's standard deviation]
  * two runs of the validation mode:                     bit-identical (images, gradients, weights after 3 steps)
"""
import ctypes
import os

import numpy as np
import pytest
import torch
import torch.nn.functional as F

pytestmark = pytest.mark.gpu

cgb = pytest.importorskip("unpaired_image_generation_b200")
from oracle import cyclegan_standin as ref  # noqa: E402
from unpaired_image_generation_b200 import _lib  # noqa: E402

GOLD = os.path.join(os.path.dirname(__file__), "golden")
NETS = ("G_AB", "G_BA", "D_A", "D_B")


def _need_gpu():
    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")


def rel(a, b):
    a, b = a.detach().double().cpu(), b.detach().double().cpu()
    return float((a - b).norm() / (b.norm() + 1e-300))


def _p(t):
    return ctypes.c_void_p(t.data_ptr()) if t is not None else None


def _dead():
    return set(ref.dead_bias_names_generator() + ref.dead_bias_names_discriminator())


# ------------------------------------------------------------------------------------------------
# single layers vs float64
# ------------------------------------------------------------------------------------------------
LAYERS = [
    # name, n, h, cin, cout, k, stride, pad, reflect, transposed, act
    ("g.res 256->256 @32 b2", 2, 32, 256, 256, 3, 1, 1, 1, 0, 0),
    ("g.stem 3->64 @64", 1, 64, 3, 64, 7, 1, 3, 1, 0, 0),
    ("g.down1 64->128 @64", 1, 64, 64, 128, 3, 2, 1, 0, 0, 0),
    ("g.up1 256->128 @16", 1, 16, 256, 128, 3, 2, 1, 0, 1, 0),
    ("g.up2 128->64 @32", 2, 32, 128, 64, 3, 2, 1, 0, 1, 0),
    ("g.head 64->3 tanh @64", 1, 64, 64, 3, 7, 1, 3, 1, 0, 2),
    ("d.conv0 3->64 leaky @64", 1, 64, 3, 64, 4, 2, 1, 0, 0, 1),
    ("d.conv2 128->256 @32", 1, 32, 128, 256, 4, 2, 1, 0, 0, 0),
    ("d.conv3 256->512 @16", 1, 16, 256, 512, 4, 1, 1, 0, 0, 0),
    ("d.conv4 512->1 @15", 1, 15, 512, 1, 4, 1, 1, 0, 0, 0),
    ("ragged 13x13", 1, 13, 64, 64, 3, 1, 1, 0, 0, 0),
]


def _torch_layer(x, w, b, stride, pad, reflect, transposed, act):
    if transposed:
        y = F.conv_transpose2d(x, w, b, stride=stride, padding=pad, output_padding=1)
    elif reflect:
        y = F.conv2d(F.pad(x, (pad,) * 4, mode="reflect"), w, b, stride=stride)
    else:
        y = F.conv2d(x, w, b, stride=stride, padding=pad)
    if act == 1:
        y = F.leaky_relu(y, 0.2)
    elif act == 2:
        y = torch.tanh(y)
    return y


@pytest.mark.parametrize("cfg", LAYERS, ids=[c[0] for c in LAYERS])
def test_fp32_conv_layer_vs_float64(cfg):
    _need_gpu()
    _, n, h, cin, cout, k, stride, pad, reflect, transposed, act = cfg
    g = torch.Generator().manual_seed(11 + len(cfg[0]))
    x32 = torch.randn(n, cin, h, h, generator=g)
    wshape = (cin, cout, k, k) if transposed else (cout, cin, k, k)
    w32 = torch.randn(wshape, generator=g) * 0.05
    b32 = torch.randn(cout, generator=g) * 0.1
    x, w, b = (t.double().requires_grad_(True) for t in (x32, w32, b32))
    y = _torch_layer(x, w, b, stride, pad, reflect, transposed, 0)
    dy32 = torch.randn(y.shape, generator=g)
    y.backward(dy32.double())
    y_act = _torch_layer(x, w, b, stride, pad, reflect, transposed, act).detach()

    lib = _lib.load()
    xd, wd, bd, dyd = x32.cuda(), w32.cuda(), b32.cuda(), dy32.cuda()
    yo = torch.empty(y.shape, device="cuda")
    dxo, dwo, dbo = torch.empty_like(xd), torch.empty_like(wd), torch.empty_like(bd)
    _lib.check(lib.cgb_conv_layer_test_f32(n, h, h, cin, cout, k, stride, pad, reflect, transposed, act, _p(xd), _p(wd),
                                           _p(bd), _p(dyd), _p(yo), _p(dxo), _p(dwo), _p(dbo)))
    assert rel(yo, y_act) < 2e-6, ("fprop", rel(yo, y_act))
    assert rel(dxo, x.grad) < 2e-6, ("dgrad", rel(dxo, x.grad))
    assert rel(dwo, w.grad) < 2e-6, ("wgrad", rel(dwo, w.grad))
    assert rel(dbo, b.grad) < 2e-6, ("bias grad", rel(dbo, b.grad))


@pytest.mark.parametrize("act,residual,shape", [(3, False, (2, 64, 24)), (1, False, (1, 512, 15)), (0, True, (2, 256, 16))])
def test_fp32_instance_norm_vs_float64(act, residual, shape):
    _need_gpu()
    lib = _lib.load()
    g = torch.Generator().manual_seed(5)
    n, c, h = shape
    y32 = torch.randn(n, c, h, h, generator=g) * 2 + 0.5
    r32 = torch.randn(n, c, h, h, generator=g) if residual else None
    da32 = torch.randn(n, c, h, h, generator=g)
    y = y32.double().requires_grad_(True)
    out = F.instance_norm(y, eps=1e-5)
    out = F.relu(out) if act == 3 else (F.leaky_relu(out, 0.2) if act == 1 else out)
    if residual:
        out = out + r32.double()
    out.backward(da32.double())
    yd, dad = y32.cuda(), da32.cuda()
    rd = r32.cuda() if residual else None
    o, dyo = torch.empty_like(yd), torch.empty_like(yd)
    _lib.check(lib.cgb_instnorm_test_f32(n, c, h, h, act, _p(yd), _p(rd), _p(dad), _p(o), _p(dyo)))
    assert rel(o, out) < 2e-6, rel(o, out)
    assert rel(dyo, y.grad) < 5e-6, rel(dyo, y.grad)


# ------------------------------------------------------------------------------------------------
# the full step at 256x256, batch 1 (BASELINE.json configs[1]) against the live fp32 stand-in
# ------------------------------------------------------------------------------------------------
def _mods_from(onets):
    mods = (cgb.Generator(), cgb.Generator(), cgb.Discriminator(), cgb.Discriminator())
    for m, o in zip(mods, onets):
        m.load_state_dict(o.state_dict())
    return mods


@pytest.fixture(scope="module")
def step256():
    _need_gpu()
    torch.set_num_threads(os.cpu_count() or 1)
    onets = ref.build_models(seed=0)
    real_A, real_B = ref.synthetic_pair(1, 256, seed=1234)
    tr = cgb.CycleGANTrainer(*_mods_from(onets), precision="fp32")
    imgs = {k: v.cpu() for k, v in tr.forward_only(real_A.cuda(), real_B.cuda()).items()}
    losses = tr.backward_only(real_A.cuda(), real_B.cuda())
    grads = {f"{nm}.{k}": v.detach().float().cpu().clone() for nm in NETS for k, v in tr.grads(nm).items()}
    otr = ref.CycleGANTrainer(*onets)
    olosses = otr.backward_only(real_A, real_B)
    oimgs = dict(otr.last_images)
    ograds = {f"{nm}.{k}": p.grad.detach().clone() for nm, net in zip(NETS, onets) for k, p in net.named_parameters()}
    # one optimiser step on both sides (the stand-in steps from the gradients it already holds)
    otr.opt_G.step()
    otr.opt_D.step()
    tr.train_step(real_A.cuda(), real_B.cuda())
    weights = {f"{nm}.{k}": p.detach().float().cpu().clone()
               for nm, m in zip(NETS, (tr.G_AB, tr.G_BA, tr.D_A, tr.D_B)) for k, p in m.named_parameters()}
    oweights = {f"{nm}.{k}": p.detach().clone() for nm, net in zip(NETS, onets) for k, p in net.named_parameters()}
    return dict(imgs=imgs, losses=losses, grads=grads, oimgs=oimgs, olosses=olosses, ograds=ograds, weights=weights,
                oweights=oweights, real_A=real_A, real_B=real_B)


def test_fp32_mode_losses_256(step256):
    s = step256
    for k in ref.CycleGANTrainer.LOSS_KEYS:
        err = abs(s["losses"][k] - s["olosses"][k]) / abs(s["olosses"][k])
        assert err < 1e-5, (k, s["losses"][k], s["olosses"][k], err)


def test_fp32_mode_images_256(step256):
    s = step256
    gold = np.load(os.path.join(GOLD, "standin_fp64_256.npz"))
    report = {}
    for k in ("fake_B", "rec_A", "fake_A", "rec_B", "idt_A", "idt_B"):
        e32 = rel(s["imgs"][k], s["oimgs"][k])
        want64 = torch.from_numpy(gold[f"img.{k}"])
        e64 = rel(s["imgs"][k][0, :, ::4, ::4], want64)
        report[k] = (e32, e64)
        assert e32 < 1e-5, (k, "vs fp32 stand-in", e32)
        assert e64 < 1e-5, (k, "vs float64 golden", e64)
    print("fp32 mode images (vs fp32 stand-in, vs float64 golden):", report)


def test_fp32_mode_gradients_256(step256):
    s = step256
    gold = np.load(os.path.join(GOLD, "standin_fp64_256.npz"))
    dead = set(f"{nm}.{k}" for nm in NETS for k in _dead())
    worst32, worst64 = (0.0, ""), (0.0, "")
    for name, g in s["grads"].items():
        if name in dead:
            assert float(g.abs().max()) == 0.0, name  # dead biases are skipped exactly
            continue
        e32 = rel(g, s["ograds"][name])
        e64 = rel(g.flatten()[::997], torch.from_numpy(gold[f"grad.{name}"]))
        worst32, worst64 = max(worst32, (e32, name)), max(worst64, (e64, name))
        assert e32 < 1e-2, (name, "vs fp32 stand-in", e32)
        assert e64 < 1e-2, (name, "vs float64 golden", e64)
        n64 = float(gold[f"gradnorm.{name}"])
        assert abs(float(g.double().norm()) - n64) / n64 < 2e-3, (name, "norm vs float64")
    # the layers next to the losses see almost no flipped decisions: tight
    for name in ("D_A.conv4.weight", "D_B.conv4.weight", "D_A.conv4.bias", "D_B.conv4.bias"):
        assert rel(s["grads"][name], s["ograds"][name]) < 2e-5, (name, rel(s["grads"][name], s["ograds"][name]))
    print("fp32 mode gradients: worst vs fp32 stand-in", worst32, "worst vs float64", worst64)


def test_fp32_mode_weights_after_one_step_256(step256):
    s = step256
    dead = set(f"{nm}.{k}" for nm in NETS for k in _dead())
    worst = (0.0, "")
    for name, w in s["weights"].items():
        if name in dead:
            continue
        if name.endswith("bias"):
            assert float((w - s["oweights"][name]).abs().max()) <= 2 * 2e-4 + 1e-7, name
            continue
        e = rel(w, s["oweights"][name])
        worst = max(worst, (e, name))
        assert e < 2e-3, (name, e)
    print("fp32 mode weights after one step: worst", worst)


# ------------------------------------------------------------------------------------------------
# determinism, checkpoint / resume
# ------------------------------------------------------------------------------------------------
def _flat_state(tr):
    torch.cuda.synchronize()
    return [t.detach().clone() for t in tr.engine.params + tr.engine.exp_avg + tr.engine.exp_avg_sq]


def test_fp32_mode_is_bit_identical_run_to_run():
    """two independent engines, three steps each (eager first call, captured graph with 8 parallel lanes after that)"""
    _need_gpu()
    real_A, real_B = ref.synthetic_pair(2, 64, seed=3)
    runs = []
    for _ in range(2):
        mods = (cgb.Generator(seed=1), cgb.Generator(seed=2), cgb.Discriminator(seed=3), cgb.Discriminator(seed=4))
        tr = cgb.CycleGANTrainer(*mods, precision="fp32")
        losses = [tr.train_step(real_A.cuda(), real_B.cuda()) for _ in range(3)]
        imgs = [tr.engine.get_image(k).clone() for k in ("fake_B", "rec_A", "idt_B")]
        runs.append((losses, imgs, _flat_state(tr), [g.clone() for g in tr.engine.grads]))
    assert runs[0][0] == runs[1][0]
    for a, b in zip(runs[0][1] + runs[0][2] + runs[0][3], runs[1][1] + runs[1][2] + runs[1][3]):
        assert torch.equal(a, b)


def test_checkpoint_resume_reproduces_step_4():
    """(params, exp_avg, exp_avg_sq, step) saved after 3 steps and restored into a fresh
    trainer: step 4 equals the uninterrupted run bit for bit (validation mode is deterministic)"""
    _need_gpu()
    real_A, real_B = ref.synthetic_pair(1, 64, seed=9)
    mk = lambda: (cgb.Generator(seed=1), cgb.Generator(seed=2), cgb.Discriminator(seed=3), cgb.Discriminator(seed=4))
    tr = cgb.CycleGANTrainer(*mk(), precision="fp32")
    for _ in range(3):
        tr.train_step(real_A.cuda(), real_B.cuda())
    ckpt = tr.state_dict()
    assert ckpt["optimizer"]["step"] == [3, 3]
    want_losses = tr.train_step(real_A.cuda(), real_B.cuda())
    want = _flat_state(tr)
    fresh = cgb.CycleGANTrainer(cgb.Generator(seed=7), cgb.Generator(seed=8), cgb.Discriminator(seed=9),
                                cgb.Discriminator(seed=10), precision="fp32")
    fresh.load_state_dict(ckpt, like=real_A)
    assert fresh.engine.step_count(0) == 3 and fresh.engine.step_count(1) == 3
    got_losses = fresh.train_step(real_A.cuda(), real_B.cuda())
    assert got_losses == want_losses
    for a, b in zip(_flat_state(fresh), want):
        assert torch.equal(a, b)
    assert fresh.engine.step_count(0) == 4
    # the optimiser state matters: the same weights WITHOUT it give a different fourth step (Adam restarts at t = 1)
    cold = cgb.CycleGANTrainer(*mk(), precision="fp32")
    cold._ensure_engine(real_A)
    for name in NETS:
        getattr(cold, name).load_state_dict(ckpt[name])
    cold.train_step(real_A.cuda(), real_B.cuda())
    assert not torch.equal(_flat_state(cold)[0], want[0])


def test_checkpoint_resume_bf16_mode():
    """the same round trip through the product (bf16) path: atomics make runs differ in the last bits, so the resumed
    step is compared at the noise level of two uninterrupted runs (losses 1e-2)"""
    _need_gpu()
    real_A, real_B = ref.synthetic_pair(1, 64, seed=9)
    mk = lambda: (cgb.Generator(seed=1), cgb.Generator(seed=2), cgb.Discriminator(seed=3), cgb.Discriminator(seed=4))
    tr = cgb.CycleGANTrainer(*mk())
    for _ in range(3):
        tr.train_step(real_A.cuda(), real_B.cuda())
    ckpt = tr.state_dict()
    want = tr.train_step(real_A.cuda(), real_B.cuda())
    fresh = cgb.CycleGANTrainer(*mk())
    fresh.load_state_dict(ckpt, like=real_A)
    got = fresh.train_step(real_A.cuda(), real_B.cuda())
    for k in want:
        assert abs(got[k] - want[k]) <= 2e-2 * abs(want[k]) + 1e-4, (k, got[k], want[k])
    assert fresh.engine.step_count(0) == 4 and fresh.engine.step_count(1) == 4
