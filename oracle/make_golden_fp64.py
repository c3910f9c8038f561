"""Golden vectors for the fp32 VALIDATION MODE (test infrastructure): the stand-in run in float64.

The fp32 stand-in is itself only an approximation of the exact step: measured here (profiles/
r02_fp32_noise_floor.txt, written by this script) it deviates from its own float64 run by up to 7e-6 (relative L2) on the
reconstructed images and by up to 4e-3 on generator weight gradients (ReLU / L1-sign decisions that flip under 1e-7
perturbations change a gradient by O(sqrt(fraction flipped))).  The B200 validation mode accumulates in fp64, so it is
compared against BOTH: the live fp32 stand-in (north_star: 1e-5 on activations and losses) and these float64 vectors.

    python -m oracle.make_golden_fp64            (about one minute on 8 cores)

Writes tests/golden/standin_fp64_256.npz: strided samples of the six images, the nine losses, per-tensor gradient
norms and every 997th gradient element, all from the float64 run at 256x256, batch 1, seeds of BASELINE.md section 3.
"""
from __future__ import annotations

import os

import numpy as np
import torch

from oracle import cyclegan_standin as S

HERE = os.path.dirname(os.path.abspath(__file__))
OUT = os.path.join(HERE, "..", "tests", "golden", "standin_fp64_256.npz")
NOISE = os.path.join(HERE, "..", "profiles", "r02_fp32_noise_floor.txt")
NETS = ("G_AB", "G_BA", "D_A", "D_B")


def run(dtype, size=256, batch=1):
    nets = [n.to(dtype) for n in S.build_models(0)]
    tr = S.CycleGANTrainer(*nets)
    a, b = S.synthetic_pair(batch, size)
    losses = tr.backward_only(a.to(dtype), b.to(dtype))
    grads = {}
    for nm, net in zip(NETS, nets):
        for k, p in net.named_parameters():
            grads[f"{nm}.{k}"] = p.grad.detach().double().clone()
    imgs = {k: v.double() for k, v in tr.last_images.items()}
    return losses, imgs, grads


def main():
    torch.set_num_threads(max(1, os.cpu_count() or 1))
    L32, I32, G32 = run(torch.float32)
    L64, I64, G64 = run(torch.float64)
    rel = lambda x, y: float((x - y).norm() / (y.norm() + 1e-300))
    dead = set([f"{n}.{k}" for n in NETS[:2] for k in S.dead_bias_names_generator()] +
               [f"{n}.{k}" for n in NETS[2:] for k in S.dead_bias_names_discriminator()])
    out = {}
    lines = ["# fp32 stand-in vs its own float64 run, 256x256 batch 1 (oracle/make_golden_fp64.py): the noise floor of any",
             "# comparison against the fp32 stand-in.  relative L2 error unless stated."]
    for k in L64:
        out[f"loss.{k}"] = np.float64(L64[k])
        lines.append(f"loss {k:14s} fp32 {L32[k]:.9g} fp64 {L64[k]:.12g} rel {abs(L32[k] - L64[k]) / abs(L64[k]):.2e}")
    for k in I64:
        out[f"img.{k}"] = I64[k][0, :, ::4, ::4].numpy().astype(np.float64)
        lines.append(f"image {k:8s} rel {rel(I32[k], I64[k]):.2e}  max abs {float((I32[k] - I64[k]).abs().max()):.2e}")
    worst = []
    for k in G64:
        if k in dead:
            continue
        out[f"gradnorm.{k}"] = np.float64(G64[k].norm())
        out[f"grad.{k}"] = G64[k].flatten()[::997].numpy().astype(np.float64)
        worst.append((rel(G32[k], G64[k]), k))
    worst.sort(reverse=True)
    lines.append(f"weight gradients (live tensors: {len(worst)}): worst {worst[0][0]:.2e} ({worst[0][1]}), "
                 f"median {worst[len(worst) // 2][0]:.2e}, best {worst[-1][0]:.2e} ({worst[-1][1]})")
    for r, k in worst[:8]:
        lines.append(f"  grad {k:28s} rel {r:.2e}")
    np.savez_compressed(OUT, **out)
    with open(NOISE, "w") as f:
        f.write("\n".join(lines) + "\n")
    print("\n".join(lines))


if __name__ == "__main__":
    main()
