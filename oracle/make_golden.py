"""Generate the frozen golden vectors for the STAND-IN oracle (test infrastructure).

The reference repository holds no golden vectors (`/root/reference/README.md:1` is all
there is), so the stand-in's own seeded outputs are the pin.  Run from the repo root:

    python -m oracle.make_golden

Writes tests/golden/standin_golden.json (scalars) and tests/golden/standin_samples.npz
(strided samples of generated images / gradients).  Re-running must reproduce the
committed files to ~1e-6 (CPU fp32 summation order may differ across torch builds).
"""
from __future__ import annotations

import json
import os

import numpy as np
import torch

from oracle.cyclegan_standin import CycleGANTrainer, build_models, synthetic_pair

HERE = os.path.dirname(os.path.abspath(__file__))
OUT = os.path.join(HERE, "..", "tests", "golden")


def tensor_stats(t: torch.Tensor):
    t = t.detach().double()
    return {"sum": float(t.sum()), "abs_sum": float(t.abs().sum()), "l2": float(t.norm())}


def run_case(size: int, batch: int, emulate: bool, steps: int):
    G_AB, G_BA, D_A, D_B = build_models(seed=0)
    tr = CycleGANTrainer(G_AB, G_BA, D_A, D_B, emulate_bf16=emulate)
    real_A, real_B = synthetic_pair(batch, size, seed=1234)
    case = {"size": size, "batch": batch, "emulate_bf16": emulate}
    imgs = tr.forward_only(real_A, real_B)
    case["images"] = {k: tensor_stats(v) for k, v in imgs.items()}
    with torch.no_grad():
        case["D_A_fake_B"] = tensor_stats(D_A(imgs["fake_B"], tr.P))
    samples = {f"img_{k}": v[0, :, ::8, ::8].numpy().copy() for k, v in imgs.items()}
    # gradients without the optimizer step
    case["losses_step0"] = tr.backward_only(real_A, real_B)
    grads = {}
    for net_name, net in (("G_AB", G_AB), ("G_BA", G_BA), ("D_A", D_A), ("D_B", D_B)):
        for n, p in net.named_parameters():
            if n.endswith("weight") and (n.startswith(("stem", "down2", "res.4.conv1", "up1", "head", "conv0", "conv3", "conv4"))):
                grads[f"{net_name}.{n}"] = tensor_stats(p.grad)
    case["grads_step0"] = grads
    samples["grad_G_AB_head_w"] = G_AB.head.weight.grad.numpy().copy()
    samples["grad_D_A_conv4_w"] = D_A.conv4.weight.grad.numpy().copy()
    # real steps
    step_losses = []
    for _ in range(steps):
        step_losses.append(tr.train_step(real_A, real_B))
    case["losses_steps"] = step_losses
    case["weights_after"] = {
        "G_AB.res.4.conv1.weight": tensor_stats(G_AB.res[4].conv1.weight),
        "G_BA.head.weight": tensor_stats(G_BA.head.weight),
        "G_BA.head.bias": tensor_stats(G_BA.head.bias),
        "D_A.conv0.weight": tensor_stats(D_A.conv0.weight),
        "D_B.conv4.bias": tensor_stats(D_B.conv4.bias),
    }
    return case, samples


def main():
    torch.set_num_threads(max(1, os.cpu_count() or 1))
    os.makedirs(OUT, exist_ok=True)
    golden = {"generator": "oracle/make_golden.py", "torch": torch.__version__, "cases": {}}
    all_samples = {}
    for name, size, batch, emulate, steps in (
        ("fp32_64", 64, 1, False, 2),
        ("bf16emu_64", 64, 1, True, 2),
        ("fp32_64_b2", 64, 2, False, 1),
        ("fp32_256", 256, 1, False, 1),
    ):
        case, samples = run_case(size, batch, emulate, steps)
        golden["cases"][name] = case
        if size == 64:
            for k, v in samples.items():
                all_samples[f"{name}.{k}"] = v.astype(np.float32)
        print(name, case["losses_step0"])
    # parameter inventory
    G_AB, _, D_A, _ = build_models(seed=0)
    golden["param_counts"] = {
        "G": sum(p.numel() for p in G_AB.parameters()), "G_tensors": len(list(G_AB.parameters())),
        "D": sum(p.numel() for p in D_A.parameters()), "D_tensors": len(list(D_A.parameters())),
    }
    with open(os.path.join(OUT, "standin_golden.json"), "w") as f:
        json.dump(golden, f, indent=1, sort_keys=True)
    np.savez_compressed(os.path.join(OUT, "standin_samples.npz"), **all_samples)


if __name__ == "__main__":
    main()
