"""Step-0 loss goldens of the STAND-IN oracle at the FULL sizes of BASELINE.json configs[2] and configs[3]
(test infrastructure; the reference holds no fixtures: /root/reference/README.md:1).

    python -m oracle.make_golden_full

  fp32_256_b8   256x256, batch 8 per GPU  (configs[2])
  fp32_512_b4   512x512, batch 4 per GPU  (configs[3])

InstanceNorm is per sample and every loss is a mean over equally sized samples, so the batch losses are the
averages of the single-sample losses: each sample is run alone (forward only), which keeps the CPU memory small.
Writes tests/golden/standin_golden_full.json.
"""
from __future__ import annotations

import json
import os

import torch

from oracle.cyclegan_standin import CycleGANTrainer, build_models, synthetic_pair

HERE = os.path.dirname(os.path.abspath(__file__))
OUT = os.path.join(HERE, "..", "tests", "golden", "standin_golden_full.json")


def batch_losses(size: int, batch: int):
    tr = CycleGANTrainer(*build_models(seed=0))
    real_A, real_B = synthetic_pair(batch, size, seed=1234)
    acc = {k: 0.0 for k in CycleGANTrainer.LOSS_KEYS}
    with torch.no_grad():
        for i in range(batch):
            a, b = real_A[i:i + 1], real_B[i:i + 1]
            L, imgs = tr.compute_G_losses(a, b)
            L = {k: float(v) for k, v in L.items()}
            L["loss_D_A"] = float(tr.compute_D_loss(tr.D_A, b, imgs["fake_B"]))
            L["loss_D_B"] = float(tr.compute_D_loss(tr.D_B, a, imgs["fake_A"]))
            for k in acc:
                acc[k] += L[k] / batch
    return acc


def main():
    torch.set_num_threads(max(1, os.cpu_count() or 1))
    golden = {"generator": "oracle/make_golden_full.py", "torch": torch.__version__, "cases": {}}
    for name, size, batch in (("fp32_256_b8", 256, 8), ("fp32_512_b4", 512, 4)):
        golden["cases"][name] = {"size": size, "batch": batch, "losses_step0": batch_losses(size, batch)}
        print(name, golden["cases"][name]["losses_step0"], flush=True)
    with open(OUT, "w") as f:
        json.dump(golden, f, indent=1, sort_keys=True)


if __name__ == "__main__":
    main()
