"""CPU-only tests: the C-ABI library loads and exports what include/cyclegan_b200.h declares, the
host-side inventory / layout logic matches the stand-in, and the product fails loudly without a GPU."""
import ctypes
import json
import os
import re

import pytest
import torch

import unpaired_image_generation_b200 as cgb
from oracle import cyclegan_standin as ref
from unpaired_image_generation_b200 import _lib

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_library_exports_every_declared_symbol():
    header = open(os.path.join(ROOT, "include", "cyclegan_b200.h")).read()
    header = re.sub(r"/\*.*?\*/", "", header, flags=re.S)
    declared = set(re.findall(r"\b(cgb_[a-z0-9_]+)\s*\(", header))
    assert len(declared) >= 25
    lib = _lib.load()
    for name in declared:
        assert hasattr(lib, name), f"{name} declared in the header but not exported"
    assert declared == set(_lib.SIGNATURES), declared ^ set(_lib.SIGNATURES)
    assert lib.cgb_version() >= 100


@pytest.mark.parametrize("n_blocks", [9, 2])
def test_inventory_matches_standin_state_dict(n_blocks):
    inv = cgb.describe(1, 64, n_blocks)
    oG = ref.Generator(n_blocks=n_blocks)
    oD = ref.Discriminator()
    for net, o in ((0, oG), (1, oG), (2, oD), (3, oD)):
        names = [(n, tuple(p.shape)) for n, p in o.named_parameters()]
        got = [(i.name, tuple(i.torch_shape)) for i in inv[net]]
        assert got == names
    if n_blocks == 9:
        assert sum(i.numel for i in inv[0]) == 11_378_179
        assert sum(i.numel for i in inv[2]) == 2_764_737
    # offsets are disjoint, ordered, and inside the flat buffers
    for group, nets in ((0, (0, 1)), (1, (2, 3))):
        end = 0
        for net in nets:
            for i in inv[net]:
                assert i.offset >= end and i.offset % 4 == 0
                end = i.offset + i.numel
        assert end <= inv["group_numel"][group]


def test_param_view_layout_roundtrip():
    inv = cgb.describe(1, 64, 1)
    flat = torch.zeros(inv["group_numel"][0])
    conv = next(i for i in inv[0] if i.name == "down1.weight")
    convT = next(i for i in inv[0] if i.name == "up1.weight")
    w = torch.randn(conv.torch_shape)
    conv.view(flat).copy_(w)
    seg = flat[conv.offset:conv.offset + conv.numel].view(conv.cout, conv.k * conv.k, conv.cin)
    assert torch.equal(seg[5, 1 * 3 + 2, 7], w[5, 7, 1, 2])          # [cout][tap][cin] <- OIHW
    wt = torch.randn(convT.torch_shape)
    convT.view(flat).copy_(wt)
    seg = flat[convT.offset:convT.offset + convT.numel].view(convT.cout, 9, convT.cin)
    assert torch.equal(seg[3, 2 * 3 + 0, 11], wt[11, 3, 2, 0])        # [cout][tap][cin] <- IOHW
    assert torch.equal(convT.view(flat), wt)


def test_engine_create_validates_config():
    lib = _lib.load()
    h = ctypes.c_void_p()
    bad = _lib.CgbConfig(1, 100, 9, 10, 10, 0.5, 2e-4, 0.5, 0.999, 1e-8)  # 100 is not a multiple of 8
    assert lib.cgb_engine_create(ctypes.byref(bad), ctypes.byref(h)) != 0
    assert b"multiple of 8" in lib.cgb_last_error()
    ok = _lib.CgbConfig(2, 256, 9, 10, 10, 0.5, 2e-4, 0.5, 0.999, 1e-8)
    assert lib.cgb_engine_create(ctypes.byref(ok), ctypes.byref(h)) == 0
    ws = lib.cgb_workspace_bytes(h)
    assert 1 << 30 < ws < 8 << 30
    # unbound engine refuses to run
    assert lib.cgb_forward_cycle(h, None) != 0
    assert b"not bound" in lib.cgb_last_error()
    lib.cgb_engine_destroy(h)


@pytest.mark.skipif(torch.cuda.is_available(), reason="checks the no-GPU failure mode")
def test_no_cpu_fallback():
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        cgb.StepEngine(1, 64)
    G = cgb.Generator(seed=0)  # parameter container works on CPU ...
    assert len(list(G.parameters())) == 48
    with pytest.raises(RuntimeError, match="no CPU fallback"):  # ... compute does not
        G(torch.zeros(1, 3, 64, 64))


def test_modules_state_dict_roundtrip_with_standin():
    oG = ref.build_models(seed=0)[0]
    G = cgb.Generator(device="cpu")
    G.load_state_dict(oG.state_dict())
    sd = G.state_dict()
    for k, v in oG.state_dict().items():
        assert torch.equal(sd[k].cpu(), v)
    with pytest.raises(KeyError):
        G.load_state_dict({"stem.weight": torch.zeros(64, 3, 7, 7)})
    oG2 = ref.Generator()
    oG2.load_state_dict({k: v.cpu() for k, v in sd.items()})  # and back into the stand-in


def test_product_never_imports_the_oracle():
    pkg = os.path.join(ROOT, "unpaired_image_generation_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cc", ".h", ".cuh")):
                src = open(os.path.join(dirpath, f)).read()
                assert "import oracle" not in src and "from oracle" not in src, f


def test_inference_engine_workspace_is_one_forward_pass():
    """cgb_engine_create_ex(CGB_FLAG_INFERENCE): host-only creation; the workspace drops the training passes"""
    import ctypes
    from unpaired_image_generation_b200 import _lib
    lib = _lib.load()
    sizes = {}
    for flags in (0, 1):
        cfg = _lib.CgbConfig(4, 256, 9, 10.0, 10.0, 0.5, 2e-4, 0.5, 0.999, 1e-8)
        h = ctypes.c_void_p()
        _lib.check(lib.cgb_engine_create_ex(ctypes.byref(cfg), flags, ctypes.byref(h)))
        sizes[flags] = lib.cgb_workspace_bytes(h)
        assert lib.cgb_group_numel(h, 0) == 22_756_360  # same parameter inventory either way
        lib.cgb_engine_destroy(h)
    assert 0 < sizes[1] < 0.25 * sizes[0], sizes


def test_paired_schedule_workspace(monkeypatch):
    """CGB_PAIR=1 re-plans the workspace (2N-image passes instead of the identity passes): host-only check"""
    import ctypes
    from unpaired_image_generation_b200 import _lib
    lib = _lib.load()
    out = {}
    for pair in ("0", "1"):
        monkeypatch.setenv("CGB_PAIR", pair)
        cfg = _lib.CgbConfig(2, 128, 9, 10.0, 10.0, 0.5, 2e-4, 0.5, 0.999, 1e-8)
        h = ctypes.c_void_p()
        _lib.check(lib.cgb_engine_create(ctypes.byref(cfg), ctypes.byref(h)))
        out[pair] = lib.cgb_workspace_bytes(h)
        lib.cgb_engine_destroy(h)
    assert 0.8 * out["0"] < out["1"] < 1.35 * out["0"], out


def test_bench_reference_arm_prints_one_contract_line():
    """bench.py --impl reference: ONE JSON line on stdout with the contract's keys (the stand-in timed on the host)"""
    import json
    import subprocess
    import sys
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    out = subprocess.run([sys.executable, os.path.join(root, "bench.py"), "--impl", "reference", "--steps", "1",
                          "--warmup", "1", "--size", "64"], capture_output=True, text=True, timeout=300, cwd=root)
    assert out.returncode == 0, out.stderr[-500:]
    lines = [l for l in out.stdout.splitlines() if l.strip()]
    assert len(lines) == 1, lines
    d = json.loads(lines[0])
    for k in ("impl", "metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better", "scaling",
              "vs_baseline", "dtype", "data", "config", "cpu_baseline", "e2e"):
        assert k in d, k
    assert d["impl"] == "reference" and d["higher_is_better"] is True and d["vs_baseline"] is None
    assert d["cpu_baseline"]["kind"] == "port" and d["cpu_baseline"]["cores"] >= 1 and d["cpu_baseline"]["value"] == d["value"]
    assert d["e2e"]["h2d_bytes_per_step"] == 0 and d["e2e"]["d2h_bytes_per_step"] == 0 and d["e2e"]["value"] == d["value"]
    assert "workload" in d["config"] and "model" not in d["config"]


def test_bench_stall_watchdog_leaves_the_process():
    """bench.py must not sit on a stalled device until the caller's time limit: the watchdog reports the phase and exits"""
    import subprocess
    import sys
    code = "import time, bench\nd = bench.StallWatchdog(1.0)\nd.beat('unit test')\ntime.sleep(30)\n"
    r = subprocess.run([sys.executable, "-c", code], capture_output=True, text=True, cwd=ROOT, timeout=60)
    assert r.returncode == 3
    assert "no progress" in json.loads(r.stdout.strip().splitlines()[-1])["error"]
