"""ctypes binding of libwavenet_b200.so (include/wavenet_b200.h).

There is no CPU or PyTorch fallback behind this module: if the shared library is missing, or
no sm_100 device is visible, the first call that needs the GPU raises.
"""
from __future__ import annotations

import ctypes as C
import os
import threading

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libwavenet_b200.so")

WN_OK, WN_ERR_INVALID, WN_ERR_CUDA, WN_ERR_UNSUPPORTED, WN_ERR_SHAPE = 0, -1, -2, -3, -4
MODE_FP32, MODE_BF16 = 0, 1
ROWS_REFERENCE, ROWS_CORRECTED = 0, 1
PUSH_OUTPUT, PUSH_INPUT = 0, 1
MODES = {"fp32": MODE_FP32, "bf16": MODE_BF16}
MODE_NAMES = ("auto", "fp32", "bf16")           # "auto": bf16 tensor-core kernels when the shape has them, else fp32 (with a warning)
ROWS = {"reference": ROWS_REFERENCE, "corrected": ROWS_CORRECTED}
PUSH = {"output": PUSH_OUTPUT, "input": PUSH_INPUT}


class wn_config(C.Structure):
    _fields_ = [("n_layers", C.c_int32), ("dilations", C.POINTER(C.c_int32)),
                ("residual_channels", C.c_int32), ("dilation_channels", C.c_int32),
                ("skip_channels", C.c_int32), ("quantization_channels", C.c_int32),
                ("use_bias", C.c_int32), ("filter_width", C.c_int32)]


class wn_gen_cond(C.Structure):
    _fields_ = [("d_fg", C.c_void_p), ("d_head", C.c_void_p), ("frames", C.c_int32), ("total_len", C.c_int32),
                ("gate_first", C.c_int32)]


class wn_ae_config(C.Structure):
    _fields_ = [("n_layers", C.c_int32), ("dilations", C.POINTER(C.c_int32)), ("quantization_channel", C.c_int32),
                ("en_residual_channel", C.c_int32), ("en_dilation_channel", C.c_int32),
                ("en_bottleneck_width", C.c_int32), ("en_pool_kernel_size", C.c_int32),
                ("de_residual_channel", C.c_int32), ("de_dilation_channel", C.c_int32), ("de_skip_channel", C.c_int32),
                ("use_bias", C.c_int32), ("filter_width", C.c_int32), ("mode", C.c_int32)]


_p, _i32, _i64, _f, _sz = C.c_void_p, C.c_int32, C.c_int64, C.c_float, C.c_size_t
_psz = C.POINTER(C.c_size_t)

# name -> (restype, argtypes); this table is also what tests/test_abi.py checks against the header
SIGNATURES = {
    "wn_version": (C.c_int, []),
    "wn_last_error": (C.c_char_p, []),
    "wn_init": (C.c_int, [C.c_int]),
    "wn_mulaw_encode": (C.c_int, [_p, _i64, _i32, _p, _p, _p]),
    "wn_mulaw_decode": (C.c_int, [_p, _i64, _i32, _p, _p, _p]),
    "wn_onehot_encode": (C.c_int, [_p, _i32, _i32, _i32, _i32, _p, _p]),
    "wn_model_create": (C.c_int, [C.POINTER(wn_config), C.POINTER(_p)]),
    "wn_model_destroy": (C.c_int, [_p]),
    "wn_model_param_count": (C.c_int64, [_p]),
    "wn_model_receptive_field": (C.c_int32, [_p]),
    "wn_model_layer_offset": (C.c_int64, [_p, _i32]),
    "wn_backward_set_split": (C.c_int, [_p, _i32, _p]),
    "wn_model_supports": (C.c_int32, [_p, _i32]),
    "wn_packed_bytes": (C.c_int, [_p, _i32, _psz]),
    "wn_pack_weights": (C.c_int, [_p, _i32, _p, _p, _p]),
    "wn_workspace_bytes": (C.c_int, [_p, _i32, _i32, _i32, _psz]),
    "wn_forward": (C.c_int, [_p, _i32, _i32, _i32, _p, _p, _p, _p, _p, _p]),
    "wn_backward": (C.c_int, [_p, _i32, _i32, _i32, _p, _p, _p, _p, _p, _p, _p]),
    "wn_softmax_fwd": (C.c_int, [_p, _i32, _i32, _i32, _i32, _p, _p]),
    "wn_softmax_bwd": (C.c_int, [_p, _p, _i32, _i32, _i32, _i32, _p, _p]),
    "wn_loss_scratch_bytes": (C.c_int, [_i32, _i32, _psz]),
    "wn_loss_fwd_bwd": (C.c_int, [_p, _p, _i32, _i32, _i32, _i32, _f, _p, _p, _p, _p]),
    "wn_adam_step": (C.c_int, [_p, _p, _p, _p, _i64, _f, _f, _f, _f, _i32, _p]),
    "wn_adam_step_dev": (C.c_int, [_p, _p, _p, _p, _i64, _f, _f, _f, _f, _p, _p]),
    "wn_sgd_step": (C.c_int, [_p, _p, _p, _i64, _f, _f, _i32, _p]),
    "wn_rmsprop_step": (C.c_int, [_p, _p, _p, _p, _i64, _f, _f, _f, _f, _p]),
    "wn_gen_state_bytes": (C.c_int, [_p, _i32, _i32, _psz]),
    "wn_gen_prime": (C.c_int, [_p, _i32, _i32, _p, _p, _p, _p, _sz, _p, _p, _p, _p]),
    "wn_gen_steps": (C.c_int, [_p, _i32, _i32, _i32, _i32, _p, _p, _p, _p, _p, _p, _p]),
    "wn_gen_export": (C.c_int, [_p, _i32, _i32, _p, _p, _p, _p]),
    "wn_gen_import": (C.c_int, [_p, _i32, _i32, _p, _p, _p, _p]),
    "wn_selftest_umma": (C.c_int, [_p, _i32, _p]),
    "wn_test_impose_relu_masks": (C.c_int, [_p, _i32, _i32, _p, _p, _p]),
    "wn_ae_create": (C.c_int, [_p, C.POINTER(_p)]),
    "wn_ae_destroy": (C.c_int, [_p]),
    "wn_ae_param_count": (C.c_int64, [_p]),
    "wn_ae_cond_param_count": (C.c_int64, [_p]),
    "wn_ae_receptive_field": (C.c_int32, [_p]),
    "wn_ae_workspace_bytes": (C.c_int, [_p, _i32, _i32, _psz]),
    "wn_ae_forward": (C.c_int, [_p, _i32, _i32, _p, _p, _p, _p, _p, _p, _p, _p]),
    "wn_set_conditioning": (C.c_int, [_p, _p]),
    "wn_ae_cond_tables": (C.c_int, [_p, _i32, _i32, _p, _p, _p, _p, _p, _p]),
    "wn_ae_train_workspace_bytes": (C.c_int, [_p, _i32, _i32, _psz]),
    "wn_ae_forward_train": (C.c_int, [_p, _i32, _i32, _p, _p, _p, _p, _p, _p, _p, _p]),
    "wn_ae_backward": (C.c_int, [_p, _i32, _i32, _p, _p, _p, _p, _p, _p, _p, _p, _p]),
    "wn_launch_count": (C.c_uint64, []),
    "wn_profile_enable": (C.c_int, [_i32]),
    "wn_profile_report": (C.c_int, [C.c_char_p, _sz]),
    "wn_bench_l2_read": (C.c_int, [_p, _i64, _i32, _i32, _p, _p]),
    "wn_debug_ts": (C.c_int, [_p, _i32]),
}

_lib = None
_lock = threading.Lock()
_inited_device = None


class WavenetB200Error(RuntimeError):
    pass


def load() -> C.CDLL:
    """dlopen the library and bind every symbol of the header. No GPU needed for this."""
    global _lib
    with _lock:
        if _lib is not None:
            return _lib
        if not os.path.exists(LIB_PATH):
            raise WavenetB200Error(
                f"{LIB_PATH} is missing: build it with `python -m music_b200.build` "
                "(music_b200 has no CPU / PyTorch fallback)")
        lib = C.CDLL(LIB_PATH)
        for name, (res, args) in SIGNATURES.items():
            fn = getattr(lib, name)
            fn.restype = res
            fn.argtypes = args
        _lib = lib
        return lib


def last_error() -> str:
    return load().wn_last_error().decode("utf-8", "replace")


def check(status: int) -> None:
    if status == WN_OK:
        return
    msg = last_error()
    if status == WN_ERR_SHAPE:
        raise ValueError(msg)          # the reference raises ValueError("wave sample not long enough")
    raise WavenetB200Error(f"libwavenet_b200 status {status}: {msg}")


def init(device_index: int = 0) -> C.CDLL:
    """wn_init on the given CUDA device (once per process per device switch)."""
    global _inited_device
    lib = load()
    if _inited_device != device_index:
        check(lib.wn_init(int(device_index)))
        _inited_device = device_index
    return lib


def profile_report():
    """[(kernel name, launches, total ms)] since wn_profile_enable(1), longest first."""
    lib = load()
    buf = C.create_string_buffer(1 << 16)
    check(lib.wn_profile_report(buf, len(buf)))
    out = []
    for line in buf.value.decode().splitlines():
        name, cnt, ms = line.rsplit(" ", 2)
        out.append((name, int(cnt), float(ms)))
    return out


def ptr(t) -> C.c_void_p:
    """Device (or host) pointer of a torch tensor, None -> NULL."""
    if t is None:
        return C.c_void_p(0)
    return C.c_void_p(t.data_ptr())


def stream_ptr():
    import torch
    return C.c_void_p(torch.cuda.current_stream().cuda_stream)


def make_model(dilations, R, D, S, Q, use_bias, filter_width=2):
    lib = load()
    arr = (C.c_int32 * len(dilations))(*[int(d) for d in dilations])
    cfg = wn_config(len(dilations), arr, int(R), int(D), int(S), int(Q), int(bool(use_bias)), int(filter_width))
    h = C.c_void_p()
    check(lib.wn_model_create(C.byref(cfg), C.byref(h)))
    return h
