"""two 64x64 train steps through the public API (run under compute-sanitizer by scripts/gpu_sanitize.sh)"""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import unpaired_image_generation_b200 as cgb  # noqa: E402

precision = sys.argv[1] if len(sys.argv) > 1 else "bf16"
mods = (cgb.Generator(seed=1, n_blocks=2), cgb.Generator(seed=2, n_blocks=2), cgb.Discriminator(seed=3), cgb.Discriminator(seed=4))
tr = cgb.CycleGANTrainer(*mods, precision=precision)
a = (torch.rand(1, 3, 64, 64) * 2 - 1).cuda()
b = (torch.rand(1, 3, 64, 64) * 2 - 1).cuda()
for _ in range(2):
    print(tr.train_step(a, b))
torch.cuda.synchronize()
print("san_step ok")
