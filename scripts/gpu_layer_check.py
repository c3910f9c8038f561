"""single-layer check of the 3-channel layers through the C ABI harness (fprop / dgrad / wgrad vs torch fp32)"""
import ctypes
import os
import sys

import torch
import torch.nn.functional as F

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from unpaired_image_generation_b200 import _lib  # noqa: E402

lib = _lib.load()
bf = lambda x: x.to(torch.bfloat16).to(torch.float32)
rel = lambda a, b: float((a.double().cpu() - b.double().cpu()).norm() / (b.double().cpu().norm() + 1e-30))
p = lambda t: ctypes.c_void_p(t.data_ptr()) if t is not None else None
CASES = [("head", 64, 3, 7, 1, 3, 1), ("stem", 3, 64, 7, 1, 3, 1), ("conv4", 512, 1, 4, 1, 1, 0), ("conv0", 3, 64, 4, 2, 1, 0)]
for name, cin, cout, k, stride, pad, reflect in CASES:
    for n, h in ((1, 32), (2, 64), (1, 128), (2, 134)):
        if name == "conv4":
            h = h // 4 + 3
        g = torch.Generator().manual_seed(n * 1000 + h)
        x = bf(torch.randn(n, cin, h, h, generator=g)).requires_grad_(True)
        w = bf(torch.randn(cout, cin, k, k, generator=g) * 0.05).requires_grad_(True)
        b = (torch.randn(cout, generator=g) * 0.1).requires_grad_(True)
        xi = F.pad(x, (pad,) * 4, mode="reflect") if reflect else x
        y = F.conv2d(xi, w, b, stride=stride, padding=0 if reflect else pad)
        dy = bf(torch.randn(y.shape, generator=g))
        y.backward(dy)
        xd, wd, bd, dyd = (t.detach().contiguous().cuda() for t in (x, w, b, dy))
        yo = torch.empty_like(y.detach()).cuda()
        dxo, dwo, dbo = torch.empty_like(xd), torch.empty_like(wd), torch.empty_like(bd)
        _lib.check(lib.cgb_conv_layer_test(n, h, h, cin, cout, k, stride, pad, reflect, 0, 0, p(xd), p(wd), p(bd), p(dyd),
                                           p(yo), p(dxo), p(dwo), p(dbo)))
        print(f"{name:6s} n={n} h={h:3d}: fprop {rel(yo, y):.2e} dgrad {rel(dxo, x.grad):.2e} wgrad {rel(dwo, w.grad):.2e} "
              f"finite {bool(torch.isfinite(dwo).all())} bias {rel(dbo, b.grad):.2e}", flush=True)
