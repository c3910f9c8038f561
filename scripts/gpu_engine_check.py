"""Development check (GPU box): B200 engine vs the stand-in oracle, with per-stage diagnostics."""
import os
import sys
import time

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import unpaired_image_generation_b200 as cgb  # noqa: E402
from oracle import cyclegan_standin as ref  # noqa: E402


def rel(a, b):
    a, b = a.detach().float().cpu(), b.detach().float().cpu()
    return float((a - b).norm() / (b.norm() + 1e-30))


def main(size=64, batch=1):
    torch.set_num_threads(os.cpu_count())
    oG_AB, oG_BA, oD_A, oD_B = ref.build_models(seed=0)
    real_A, real_B = ref.synthetic_pair(batch, size, seed=1234)
    G_AB, G_BA, D_A, D_B = cgb.Generator(), cgb.Generator(), cgb.Discriminator(), cgb.Discriminator()
    for m, o in ((G_AB, oG_AB), (G_BA, oG_BA), (D_A, oD_A), (D_B, oD_B)):
        m.load_state_dict(o.state_dict())
    tr = cgb.CycleGANTrainer(G_AB, G_BA, D_A, D_B)
    otr = ref.CycleGANTrainer(oG_AB, oG_BA, oD_A, oD_B, emulate_bf16=True)
    otr32 = ref.CycleGANTrainer(*ref.build_models(seed=0))

    xa, xb = real_A.cuda(), real_B.cuda()
    # module-level forward
    y = G_AB(xa)
    cap = {}
    with torch.no_grad():
        yo = oG_AB(real_A, otr.P, capture=cap)
    print(f"[{size}] G_AB.forward vs bf16-emulated oracle: rel {rel(y, yo):.3e}")
    d = D_A(xa)
    with torch.no_grad():
        do = oD_A(real_A, otr.P)
    print(f"[{size}] D_A.forward  vs bf16-emulated oracle: rel {rel(d, do):.3e}")

    imgs = tr.forward_only(xa, xb)
    oimgs = otr.forward_only(real_A, real_B)
    oimgs32 = otr32.forward_only(real_A, real_B)
    for k in imgs:
        print(f"[{size}] {k:7s} rel vs emu {rel(imgs[k], oimgs[k]):.3e}   vs fp32 {rel(imgs[k], oimgs32[k]):.3e}"
              f"   (emu vs fp32 {rel(oimgs[k], oimgs32[k]):.3e})")

    losses = tr.backward_only(xa, xb)
    olosses = otr.backward_only(real_A, real_B)
    olosses32 = otr32.backward_only(real_A, real_B)
    for k in losses:
        print(f"[{size}] {k:13s} gpu {losses[k]:.5f}  emu {olosses[k]:.5f}  fp32 {olosses32[k]:.5f}")
    for net_name, onet in (("G_AB", oG_AB), ("G_BA", oG_BA), ("D_A", oD_A), ("D_B", oD_B)):
        g = tr.grads(net_name)
        for n, p in onet.named_parameters():
            if p.grad is None:
                continue
            a, b = g[n].detach().float().cpu(), p.grad
            if float(b.abs().max()) < 1e-5:
                continue
            cos = float((a * b).sum() / (a.norm() * b.norm() + 1e-30))
            tag = "" if rel(a, b) < 0.1 else "   <<<<<<"
            if n.endswith("weight") or rel(a, b) > 0.1:
                print(f"[{size}] grad {net_name}.{n:22s} rel {rel(a, b):.3e} cos {cos:.4f}{tag}")
    # one real step + timing
    l1 = tr.train_step(xa, xb)
    ol1 = otr.train_step(real_A, real_B)
    print(f"[{size}] step1 loss_G gpu {l1['loss_G']:.5f} emu {ol1['loss_G']:.5f}; loss_D_A gpu {l1['loss_D_A']:.5f} emu {ol1['loss_D_A']:.5f}")
    for net_name, mod, onet in (("G_AB", G_AB, oG_AB), ("D_A", D_A, oD_A)):
        worst = 0.0
        dead = set(ref.dead_bias_names_generator() + ref.dead_bias_names_discriminator())
        for (n, p), (_, q) in zip(mod.named_parameters(), onet.named_parameters()):
            if n in dead:
                continue
            worst = max(worst, rel(p, q))
        print(f"[{size}] weights after one step {net_name}: worst rel {worst:.3e}")
    for i in range(3):
        tr.train_step(xa, xb)
    torch.cuda.synchronize()
    t0 = time.time()
    n = 10
    for i in range(n):
        tr.train_step(xa, xb)
    torch.cuda.synchronize()
    dt = (time.time() - t0) / n
    print(f"[{size}] train_step {dt * 1e3:.2f} ms/step = {batch / dt:.1f} img/s; launches/step {tr.engine.launches_per_step}; "
          f"conv TFLOP/step {tr.engine.conv_flops_per_step / 1e12:.3f} -> {tr.engine.conv_flops_per_step / dt / 1e12:.1f} TFLOP/s")


if __name__ == "__main__":
    main(int(sys.argv[1]) if len(sys.argv) > 1 else 64, int(sys.argv[2]) if len(sys.argv) > 2 else 1)
