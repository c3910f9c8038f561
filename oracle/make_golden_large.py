"""Golden vectors of the STAND-IN oracle for the larger BASELINE.json configurations (test infrastructure).

    python -m oracle.make_golden_large

Cases (the reference itself holds no fixtures: /root/reference/README.md:1):
  fp32_512_b1   configs[3] geometry (512x512; batch 1 keeps the CPU run short): images, step-0 losses, one step
  fp32_256_b2   configs[2] geometry (256x256, batch > 1): step-0 losses
  gen_512_b2, gen_1024_b1   configs[4], generator-only inference: statistics and a strided sample of G_AB(x)
Writes tests/golden/standin_golden_large.json and tests/golden/standin_samples_large.npz.
"""
from __future__ import annotations

import json
import os

import numpy as np
import torch

from oracle.cyclegan_standin import CycleGANTrainer, build_models, synthetic_pair
from oracle.make_golden import tensor_stats

HERE = os.path.dirname(os.path.abspath(__file__))
OUT = os.path.join(HERE, "..", "tests", "golden")


def train_case(size: int, batch: int, steps: int):
    G_AB, G_BA, D_A, D_B = build_models(seed=0)
    tr = CycleGANTrainer(G_AB, G_BA, D_A, D_B)
    real_A, real_B = synthetic_pair(batch, size, seed=1234)
    case = {"size": size, "batch": batch}
    imgs = tr.forward_only(real_A, real_B)
    case["images"] = {k: tensor_stats(v) for k, v in imgs.items()}
    case["losses_step0"] = tr.backward_only(real_A, real_B)
    case["losses_steps"] = [tr.train_step(real_A, real_B) for _ in range(steps)]
    samples = {f"img_{k}": v[0, :, ::16, ::16].numpy().copy() for k, v in imgs.items()}
    return case, samples


def gen_case(size: int, batch: int):
    G_AB, _, _, _ = build_models(seed=0)
    x, _ = synthetic_pair(batch, size, seed=4321)
    with torch.no_grad():
        y = G_AB(x)
    stride = size // 32
    return {"size": size, "batch": batch, "out": tensor_stats(y)}, {"y": y[:, :, ::stride, ::stride].numpy().copy()}


def main():
    torch.set_num_threads(max(1, os.cpu_count() or 1))
    golden = {"generator": "oracle/make_golden_large.py", "torch": torch.__version__, "cases": {}}
    all_samples = {}
    for name, size, batch, steps in (("fp32_512_b1", 512, 1, 1), ("fp32_256_b2", 256, 2, 0)):
        case, samples = train_case(size, batch, steps)
        golden["cases"][name] = case
        for k, v in samples.items():
            all_samples[f"{name}.{k}"] = v.astype(np.float32)
        print(name, case["losses_step0"], flush=True)
    for name, size, batch in (("gen_512_b2", 512, 2), ("gen_1024_b1", 1024, 1)):
        case, samples = gen_case(size, batch)
        golden["cases"][name] = case
        for k, v in samples.items():
            all_samples[f"{name}.{k}"] = v.astype(np.float32)
        print(name, case["out"], flush=True)
    with open(os.path.join(OUT, "standin_golden_large.json"), "w") as f:
        json.dump(golden, f, indent=1, sort_keys=True)
    np.savez_compressed(os.path.join(OUT, "standin_samples_large.npz"), **all_samples)


if __name__ == "__main__":
    main()
