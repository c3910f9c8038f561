"""ctypes binding of libcyclegan_b200.so (C ABI declared in include/cyclegan_b200.h).

There is NO fallback: if the shared library is missing the import of any product entry point
raises, with the build command in the message.
"""
from __future__ import annotations

import ctypes
import os
from ctypes import (POINTER, Structure, c_char, c_char_p, c_double, c_float, c_int, c_longlong, c_void_p)

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libcyclegan_b200.so")


class CgbConfig(Structure):
    _fields_ = [("batch", c_int), ("size", c_int), ("n_blocks", c_int), ("lambda_A", c_float),
                ("lambda_B", c_float), ("lambda_idt", c_float), ("lr", c_float), ("beta1", c_float),
                ("beta2", c_float), ("eps", c_float)]


class CgbParamInfo(Structure):
    _fields_ = [("name", c_char * 64), ("is_bias", c_int), ("transposed", c_int), ("cout", c_int),
                ("cin", c_int), ("k", c_int), ("offset", c_longlong), ("numel", c_longlong)]


# every symbol include/cyclegan_b200.h declares: name -> (restype, argtypes)
_P = c_void_p
SIGNATURES = {
    "cgb_last_error": (c_char_p, []),
    "cgb_version": (c_int, []),
    "cgb_engine_create": (c_int, [POINTER(CgbConfig), POINTER(_P)]),
    "cgb_engine_create_ex": (c_int, [POINTER(CgbConfig), c_int, POINTER(_P)]),
    "cgb_engine_destroy": (None, [_P]),
    "cgb_num_params": (c_int, [_P, c_int]),
    "cgb_param_info": (c_int, [_P, c_int, c_int, POINTER(CgbParamInfo)]),
    "cgb_group_numel": (c_longlong, [_P, c_int]),
    "cgb_workspace_bytes": (c_longlong, [_P]),
    "cgb_engine_bind": (c_int, [_P] + [_P] * 8 + [_P, c_longlong]),
    "cgb_refresh_weights": (c_int, [_P, c_int, _P]),
    "cgb_set_grad_scale": (c_int, [_P, c_float]),
    "cgb_set_step_count": (c_int, [_P, c_int, c_int]),
    "cgb_get_step_count": (c_int, [_P, c_int, POINTER(c_int)]),
    "cgb_set_lr": (c_int, [_P, c_int, c_float, _P]),
    "cgb_engine_set_image_pool": (c_int, [_P, c_int]),
    "cgb_set_pool_decisions": (c_int, [_P, _P, _P]),
    "cgb_stage_inputs_u8": (c_int, [_P, _P, _P, _P]),
    "cgb_generator_forward": (c_int, [_P, c_int, _P, _P, _P]),
    "cgb_discriminator_forward": (c_int, [_P, c_int, _P, _P, _P]),
    "cgb_set_inputs": (c_int, [_P, _P, _P, _P]),
    "cgb_forward_cycle": (c_int, [_P, _P]),
    "cgb_get_image": (c_int, [_P, c_int, _P, _P]),
    "cgb_get_image_u8": (c_int, [_P, c_int, _P, _P]),
    "cgb_phase_generators": (c_int, [_P, _P]),
    "cgb_phase_discriminators": (c_int, [_P, _P]),
    "cgb_adam": (c_int, [_P, c_int, _P]),
    "cgb_adam_range": (c_int, [_P, c_int, c_longlong, c_longlong, c_int, _P]),
    "cgb_train_step": (c_int, [_P, _P]),
    "cgb_stage_inputs": (c_int, [_P, _P, _P, _P]),
    "cgb_run_segment": (c_int, [_P, c_int, _P]),
    "cgb_num_grad_buckets": (c_int, [_P]),
    "cgb_grad_bucket_info": (c_int, [_P, c_int, POINTER(c_int), POINTER(c_longlong), POINTER(c_longlong)]),
    "cgb_wait_grad_bucket": (c_int, [_P, c_int, _P]),
    "cgb_grad_bucket_layers": (c_int, [_P, c_int, POINTER(c_int), POINTER(c_int), POINTER(c_int), POINTER(c_int)]),
    "cgb_refresh_weights_layers": (c_int, [_P, c_int, c_int, c_int, _P]),
    "cgb_get_losses_host": (c_int, [_P, POINTER(c_float), _P]),
    "cgb_train_step_host": (c_int, [_P, _P, _P, POINTER(c_float), _P]),
    "cgb_launches_per_step": (c_longlong, [_P]),
    "cgb_conv_flops_per_step": (c_double, [_P]),
    "cgb_profile_kind": (c_int, [_P, c_int, c_int, _P, POINTER(c_float), POINTER(c_longlong), POINTER(c_double)]),
    "cgb_profile_timeline": (c_int, [_P, _P, c_char_p, c_int]),
    "cgb_conv_layer_test": (c_int, [c_int] * 11 + [_P] * 8),
    "cgb_instnorm_test": (c_int, [c_int] * 5 + [_P] * 5),
    "cgb_instnorm_bwd_test": (c_int, [c_int] * 7 + [_P] * 5),
    "cgb_conv_layer_test_f32": (c_int, [c_int] * 11 + [_P] * 8),
    "cgb_instnorm_test_f32": (c_int, [c_int] * 5 + [_P] * 5),
}

_lib = None


def load() -> ctypes.CDLL:
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise RuntimeError(
            f"{LIB_PATH} is missing: the B200 CUDA library has not been built. Run "
            "`python -c 'import __graft_entry__ as g; g.build()'` (or `make -C "
            "unpaired_image_generation_b200/csrc`). There is no CPU fallback.")
    lib = ctypes.CDLL(LIB_PATH)
    for name, (res, args) in SIGNATURES.items():
        fn = getattr(lib, name)  # AttributeError if the .so does not export a declared symbol
        fn.restype = res
        fn.argtypes = args
    _lib = lib
    return lib


class CgbError(RuntimeError):
    pass


def check(rc: int) -> None:
    if rc != 0:
        msg = load().cgb_last_error()
        raise CgbError(msg.decode() if msg else f"libcyclegan_b200 call failed with code {rc}")
