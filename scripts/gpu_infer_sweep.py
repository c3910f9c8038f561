"""Generator-only inference sweep (BASELINE.json configs[4]): Generator.forward through inference-only engines,
CUDA-event timed, device-resident fp32 NCHW inputs.  Prints one JSON line per (size, batch)."""
import json
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import unpaired_image_generation_b200 as cgb  # noqa: E402

G = cgb.Generator(seed=1)
FWD_GFLOP_256 = 99.103
for size, batch in ((256, 1), (256, 8), (256, 64), (512, 1), (512, 16), (1024, 1), (1024, 4)):
    x = (torch.rand(batch, 3, size, size) * 2 - 1).cuda()
    for _ in range(3):
        y = G(x)
    torch.cuda.synchronize()
    reps = 10 if batch * size * size <= 64 * 256 * 256 else 5
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        y = G(x)
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / reps
    flops = FWD_GFLOP_256 * 1e9 * batch * (size / 256) ** 2
    print(json.dumps({"size": size, "batch": batch, "ms": round(ms, 3), "images_per_s": round(batch / (ms * 1e-3), 1),
                      "conv_tflops": round(flops / (ms * 1e-3) / 1e12, 1),
                      "workspace_MiB": round(G._private[(batch, size)].workspace_bytes / 2**20)}), flush=True)
    G._private.clear()
    torch.cuda.empty_cache()
