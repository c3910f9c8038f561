"""Python handle on the C++ step engine: owns the torch-allocated flat parameter / gradient / Adam
buffers and the HBM workspace, and forwards every call to the C ABI with raw device pointers.
PyTorch is used for device memory and streams only."""
from __future__ import annotations

import ctypes
from collections import OrderedDict
from typing import Dict, List

import torch

from . import _lib

NET_G_AB, NET_G_BA, NET_D_A, NET_D_B = 0, 1, 2, 3
GROUP_G, GROUP_D = 0, 1
IMAGE_IDS = OrderedDict(fake_B=0, rec_A=1, fake_A=2, rec_B=3, idt_A=4, idt_B=5, real_A=6, real_B=7,
                        pool_fake_B=8, pool_fake_A=9)
LOSS_KEYS = ("loss_G", "loss_G_A", "loss_G_B", "loss_cycle_A", "loss_cycle_B", "loss_idt_A", "loss_idt_B",
             "loss_D_A", "loss_D_B")


def _ptr(t: torch.Tensor) -> ctypes.c_void_p:
    return ctypes.c_void_p(t.data_ptr())


def _stream() -> ctypes.c_void_p:
    return ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)


class ParamInfo:
    def __init__(self, raw: "_lib.CgbParamInfo"):
        self.name = raw.name.decode()
        self.is_bias = bool(raw.is_bias)
        self.transposed = bool(raw.transposed)
        self.cout, self.cin, self.k = raw.cout, raw.cin, raw.k
        self.offset, self.numel = raw.offset, raw.numel

    @property
    def torch_shape(self):
        if self.is_bias:
            return (self.cout,)
        return (self.cin, self.cout, self.k, self.k) if self.transposed else (self.cout, self.cin, self.k, self.k)

    def view(self, flat: torch.Tensor) -> torch.Tensor:
        """torch-layout view (no copy) of this tensor inside a flat group buffer."""
        seg = flat[self.offset:self.offset + self.numel]
        if self.is_bias:
            return seg
        v = seg.view(self.cout, self.k, self.k, self.cin)
        return v.permute(3, 0, 1, 2) if self.transposed else v.permute(0, 3, 1, 2)


def describe(batch: int, size: int, n_blocks: int = 9) -> Dict[int, List[ParamInfo]]:
    """Parameter inventory without touching the GPU (engine creation is host-only)."""
    lib = _lib.load()
    cfg = _lib.CgbConfig(batch, size, n_blocks, 10.0, 10.0, 0.5, 2e-4, 0.5, 0.999, 1e-8)
    h = ctypes.c_void_p()
    _lib.check(lib.cgb_engine_create(ctypes.byref(cfg), ctypes.byref(h)))
    try:
        out = {}
        for net in range(4):
            infos = []
            for i in range(lib.cgb_num_params(h, net)):
                raw = _lib.CgbParamInfo()
                _lib.check(lib.cgb_param_info(h, net, i, ctypes.byref(raw)))
                infos.append(ParamInfo(raw))
            out[net] = infos
        out["group_numel"] = (lib.cgb_group_numel(h, 0), lib.cgb_group_numel(h, 1))
        out["workspace_bytes"] = lib.cgb_workspace_bytes(h)
        return out
    finally:
        lib.cgb_engine_destroy(h)


class StepEngine:
    def __init__(self, batch: int, size: int, n_blocks: int = 9, lambda_A: float = 10.0, lambda_B: float = 10.0,
                 lambda_idt: float = 0.5, lr: float = 2e-4, betas=(0.5, 0.999), eps: float = 1e-8, device=None,
                 inference: bool = False, pool_size: int = 0, precision: str = "bf16"):
        if not torch.cuda.is_available():
            raise RuntimeError("unpaired_image_generation_b200 needs a CUDA device (B200, sm_100a); "
                               "there is no CPU fallback")
        self.lib = _lib.load()
        self.device = torch.device(device if device is not None else f"cuda:{torch.cuda.current_device()}")
        self.batch, self.size, self.n_blocks = batch, size, n_blocks
        cfg = _lib.CgbConfig(batch, size, n_blocks, lambda_A, lambda_B, lambda_idt, lr, betas[0], betas[1], eps)
        self._h = ctypes.c_void_p()
        # inference=True: module forwards only (CGB_FLAG_INFERENCE): the workspace holds one forward pass
        # precision="fp32": the deterministic fp32 validation mode (CGB_FLAG_FP32_VALIDATE), same programs / schedule
        if precision not in ("bf16", "fp32"):
            raise ValueError("precision must be 'bf16' (product path) or 'fp32' (validation mode)")
        self.inference = bool(inference)
        self.precision = precision
        flags = (1 if inference else 0) | (2 if precision == "fp32" else 0)
        _lib.check(self.lib.cgb_engine_create_ex(ctypes.byref(cfg), flags, ctypes.byref(self._h)))
        self.pool_size = int(pool_size)
        if self.pool_size > 0:  # image history pool: re-plans the workspace, must precede the bind
            _lib.check(self.lib.cgb_engine_set_image_pool(self._h, self.pool_size))
        self.infos: Dict[int, List[ParamInfo]] = {}
        for net in range(4):
            lst = []
            for i in range(self.lib.cgb_num_params(self._h, net)):
                raw = _lib.CgbParamInfo()
                _lib.check(self.lib.cgb_param_info(self._h, net, i, ctypes.byref(raw)))
                lst.append(ParamInfo(raw))
            self.infos[net] = lst
        with torch.cuda.device(self.device):
            numel = [self.lib.cgb_group_numel(self._h, g) for g in range(2)]
            mk = lambda n: torch.zeros(n, dtype=torch.float32, device=self.device)
            self.params = [mk(numel[0]), mk(numel[1])]
            if self.inference:
                # module forwards never touch gradients or Adam moments: bind one small dummy instead of 3 x 113 MB
                dummy = mk(4)
                self.grads = self.exp_avg = self.exp_avg_sq = [dummy, dummy]
            else:
                self.grads = [mk(numel[0]), mk(numel[1])]
                self.exp_avg = [mk(numel[0]), mk(numel[1])]
                self.exp_avg_sq = [mk(numel[0]), mk(numel[1])]
            ws_bytes = self.lib.cgb_workspace_bytes(self._h)
            self._ws_raw = torch.empty(ws_bytes + 1024, dtype=torch.uint8, device=self.device)
            off = (-self._ws_raw.data_ptr()) % 1024
            self.workspace = self._ws_raw[off:off + ws_bytes]
            torch.cuda.synchronize(self.device)
            _lib.check(self.lib.cgb_engine_bind(
                self._h, _ptr(self.params[0]), _ptr(self.grads[0]), _ptr(self.exp_avg[0]), _ptr(self.exp_avg_sq[0]),
                _ptr(self.params[1]), _ptr(self.grads[1]), _ptr(self.exp_avg[1]), _ptr(self.exp_avg_sq[1]),
                _ptr(self.workspace), ws_bytes))
        self.workspace_bytes = ws_bytes
        self._losses_host = (ctypes.c_float * 16)()

    def __del__(self):
        h = getattr(self, "_h", None)
        if h:
            try:
                torch.cuda.synchronize(self.device)
            except Exception:
                pass
            self.lib.cgb_engine_destroy(h)
            self._h = None

    # ---- parameters -----------------------------------------------------------------------------
    def _views(self, net: int, bufs) -> "OrderedDict[str, torch.Tensor]":
        flat = bufs[0 if net < 2 else 1]
        return OrderedDict((i.name, i.view(flat)) for i in self.infos[net])

    def param_views(self, net: int):
        return self._views(net, self.params)

    def grad_views(self, net: int):
        return self._views(net, self.grads)

    def refresh_weights(self, group: int):
        _lib.check(self.lib.cgb_refresh_weights(self._h, group, _stream()))

    def set_grad_scale(self, scale: float):
        _lib.check(self.lib.cgb_set_grad_scale(self._h, scale))

    # ---- optimiser state (checkpoint / resume) ----------------------------------------------------
    def step_count(self, group: int) -> int:
        out = ctypes.c_int()
        _lib.check(self.lib.cgb_get_step_count(self._h, group, ctypes.byref(out)))
        return int(out.value)

    def set_step_count(self, group: int, step: int):
        _lib.check(self.lib.cgb_set_step_count(self._h, group, int(step)))

    def optimizer_state(self) -> Dict[str, object]:
        """flat Adam state of both groups (stand-in: opt_G.state_dict() / opt_D.state_dict()): moments + step"""
        return {"exp_avg": [t.detach().clone() for t in self.exp_avg],
                "exp_avg_sq": [t.detach().clone() for t in self.exp_avg_sq],
                "step": [self.step_count(0), self.step_count(1)]}

    def load_optimizer_state(self, state: Dict[str, object]):
        for g in range(2):
            self.exp_avg[g].copy_(state["exp_avg"][g])
            self.exp_avg_sq[g].copy_(state["exp_avg_sq"][g])
            self.set_step_count(g, state["step"][g])

    # ---- modules --------------------------------------------------------------------------------
    def _check_img(self, x: torch.Tensor) -> torch.Tensor:
        if tuple(x.shape) != (self.batch, 3, self.size, self.size):
            raise ValueError(f"expected input of shape {(self.batch, 3, self.size, self.size)}, got {tuple(x.shape)}")
        return x.to(device=self.device, dtype=torch.float32).contiguous()

    def generator_forward(self, net: int, x: torch.Tensor) -> torch.Tensor:
        x = self._check_img(x)
        y = torch.empty_like(x)
        _lib.check(self.lib.cgb_generator_forward(self._h, net, _ptr(x), _ptr(y), _stream()))
        return y

    def discriminator_forward(self, net: int, x: torch.Tensor) -> torch.Tensor:
        x = self._check_img(x)
        p = self.size // 8 - 2
        y = torch.empty(self.batch, 1, p, p, dtype=torch.float32, device=self.device)
        _lib.check(self.lib.cgb_discriminator_forward(self._h, net, _ptr(x), _ptr(y), _stream()))
        return y

    # ---- step -----------------------------------------------------------------------------------
    def set_inputs(self, real_A: torch.Tensor, real_B: torch.Tensor):
        a, b = self._check_img(real_A), self._check_img(real_B)
        _lib.check(self.lib.cgb_set_inputs(self._h, _ptr(a), _ptr(b), _stream()))
        self._keep = (a, b)  # keep alive until the async copies ran

    def stage_inputs(self, real_A: torch.Tensor, real_B: torch.Tensor):
        """copy the inputs into the engine's staging buffers (the graph segments read them from there)"""
        a, b = self._check_img(real_A), self._check_img(real_B)
        _lib.check(self.lib.cgb_stage_inputs(self._h, _ptr(a), _ptr(b), _stream()))
        self._keep = (a, b)

    def stage_inputs_u8(self, real_A: torch.Tensor, real_B: torch.Tensor):
        """uint8 interleaved RGB [N, H, W, 3] inputs (device or pinned host): converted on the device with
        x = u8 / 127.5 - 1 (stand-in: from_uint8); a quarter of the bytes of the fp32 path"""
        for t in (real_A, real_B):
            if t.dtype != torch.uint8 or tuple(t.shape) != (self.batch, self.size, self.size, 3) or not t.is_contiguous():
                raise ValueError(f"expected contiguous uint8 [{self.batch}, {self.size}, {self.size}, 3], got {t.dtype} {tuple(t.shape)}")
        _lib.check(self.lib.cgb_stage_inputs_u8(self._h, _ptr(real_A), _ptr(real_B), _stream()))
        self._keep = (real_A, real_B)

    def set_pool_decisions(self, decisions: torch.Tensor):
        """int32 [2, batch, 2] (store, ret) pairs of this step's image-pool queries (device or pinned host)"""
        if decisions.dtype != torch.int32 or tuple(decisions.shape) != (2, self.batch, 2) or not decisions.is_contiguous():
            raise ValueError(f"expected contiguous int32 [2, {self.batch}, 2]")
        _lib.check(self.lib.cgb_set_pool_decisions(self._h, _ptr(decisions), _stream()))
        self._keep_dec = decisions

    def set_lr(self, lr: float, group: int = -1):
        """learning rate of one parameter group (0 generators, 1 discriminators) or of both (-1); stream-ordered,
        held in device memory: no graph re-capture"""
        for g in ((0, 1) if group < 0 else (group,)):
            _lib.check(self.lib.cgb_set_lr(self._h, g, float(lr), _stream()))

    def run_segment(self, segment: int):
        """graph-replayed part of the step: 0 whole step, 1 G phase (incl. forwards), 2 D phase, 3/4 Adam G/D"""
        _lib.check(self.lib.cgb_run_segment(self._h, segment, _stream()))

    def grad_buckets(self) -> List[tuple]:
        """(group, offset, numel) ranges of the flat gradient buffers announced by the data-parallel step (segment 6),
        indexed as the C ABI indexes them"""
        out = []
        for i in range(max(0, self.lib.cgb_num_grad_buckets(self._h))):
            g, off, n = ctypes.c_int(), ctypes.c_longlong(), ctypes.c_longlong()
            _lib.check(self.lib.cgb_grad_bucket_info(self._h, i, ctypes.byref(g), ctypes.byref(off), ctypes.byref(n)))
            out.append((int(g.value), int(off.value), int(n.value)))
        return out

    def grad_bucket_plan(self) -> List[dict]:
        """the buckets in the order they become final, each with the layer range it covers:
        dict(index, group, offset, numel, net, layer_lo, layer_hi)"""
        plan = []
        for i, (g, off, n) in enumerate(self.grad_buckets()):
            net, lo, hi, order = ctypes.c_int(), ctypes.c_int(), ctypes.c_int(), ctypes.c_int()
            _lib.check(self.lib.cgb_grad_bucket_layers(self._h, i, ctypes.byref(net), ctypes.byref(lo), ctypes.byref(hi),
                                                       ctypes.byref(order)))
            plan.append(dict(index=i, group=g, offset=off, numel=n, net=int(net.value), layer_lo=int(lo.value),
                             layer_hi=int(hi.value), order=int(order.value)))
        plan.sort(key=lambda b: b["order"])
        return plan

    def refresh_weights_layers(self, net: int, layer_lo: int, layer_hi: int):
        _lib.check(self.lib.cgb_refresh_weights_layers(self._h, net, layer_lo, layer_hi, _stream()))

    def wait_grad_bucket(self, index: int):
        """the current stream waits until bucket `index` of the most recently launched segment 6 is final"""
        _lib.check(self.lib.cgb_wait_grad_bucket(self._h, index, _stream()))

    def forward_cycle(self):
        _lib.check(self.lib.cgb_forward_cycle(self._h, _stream()))

    def get_image(self, name: str) -> torch.Tensor:
        out = torch.empty(self.batch, 3, self.size, self.size, dtype=torch.float32, device=self.device)
        _lib.check(self.lib.cgb_get_image(self._h, IMAGE_IDS[name], _ptr(out), _stream()))
        return out

    def get_image_u8(self, name: str) -> torch.Tensor:
        """the image as uint8 interleaved RGB [N, H, W, 3] (stand-in: to_uint8)"""
        out = torch.empty(self.batch, self.size, self.size, 3, dtype=torch.uint8, device=self.device)
        _lib.check(self.lib.cgb_get_image_u8(self._h, IMAGE_IDS[name], _ptr(out), _stream()))
        return out

    def phase_generators(self):
        _lib.check(self.lib.cgb_phase_generators(self._h, _stream()))

    def phase_discriminators(self):
        _lib.check(self.lib.cgb_phase_discriminators(self._h, _stream()))

    def adam(self, group: int):
        _lib.check(self.lib.cgb_adam(self._h, group, _stream()))

    def adam_range(self, group: int, offset: int, numel: int, advance_step: bool):
        """Adam on a sub-range of a group's flat buffers (no bf16 refresh); see cgb_adam_range"""
        _lib.check(self.lib.cgb_adam_range(self._h, group, offset, numel, 1 if advance_step else 0, _stream()))

    def train_step(self):
        """whole step on the current stream from the staged inputs (CUDA graph after the first call)"""
        _lib.check(self.lib.cgb_train_step(self._h, _stream()))

    def losses(self) -> Dict[str, float]:
        _lib.check(self.lib.cgb_get_losses_host(self._h, self._losses_host, _stream()))
        return {k: float(self._losses_host[i]) for i, k in enumerate(LOSS_KEYS)}

    def train_step_host(self, real_A_host: torch.Tensor, real_B_host: torch.Tensor) -> Dict[str, float]:
        """end to end from HOST tensors (pinned for async copies): H2D + step + D2H of the losses"""
        assert real_A_host.device.type == "cpu" and real_A_host.dtype == torch.float32 and real_A_host.is_contiguous()
        assert real_B_host.device.type == "cpu" and real_B_host.dtype == torch.float32 and real_B_host.is_contiguous()
        _lib.check(self.lib.cgb_train_step_host(self._h, _ptr(real_A_host), _ptr(real_B_host), self._losses_host,
                                                _stream()))
        return {k: float(self._losses_host[i]) for i, k in enumerate(LOSS_KEYS)}

    def profile_kind(self, kind: int, reps: int = 5):
        """(ms per step-equivalent, launches, algorithmic FLOPs) of one kernel class, CUDA-event timed"""
        ms, n, fl = ctypes.c_float(), ctypes.c_longlong(), ctypes.c_double()
        _lib.check(self.lib.cgb_profile_kind(self._h, kind, reps, _stream(), ctypes.byref(ms), ctypes.byref(n),
                                             ctypes.byref(fl)))
        return ms.value, n.value, fl.value

    def timeline(self) -> str:
        """pass-boundary timestamps of one graph-replayed step (development profiling)"""
        buf = ctypes.create_string_buffer(65536)
        _lib.check(self.lib.cgb_profile_timeline(self._h, _stream(), buf, 65536))
        return buf.value.decode()

    @property
    def launches_per_step(self) -> int:
        return int(self.lib.cgb_launches_per_step(self._h))

    @property
    def conv_flops_per_step(self) -> float:
        return float(self.lib.cgb_conv_flops_per_step(self._h))
