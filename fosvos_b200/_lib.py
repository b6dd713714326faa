"""ctypes binding of libfosvos_sm100.so (the C ABI in include/fosvos_b200.h).

There is no CPU fallback and no other backend: if the library has not been built
(``python -m fosvos_b200.build``) or a call fails, this module raises.
"""
from __future__ import annotations

import ctypes as C
import os
from typing import Optional, Sequence

import torch

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "lib", "libfosvos_sm100.so")

F32, BF16 = 0, 1
CONV_BIAS, CONV_RELU, CONV_MASK, CONV_ACCUMULATE = 1, 2, 4, 8
W_SIMT_FWD, W_SIMT_DGRAD, W_TC_FWD, W_TC_DGRAD = 0, 1, 2, 3

_vp, _i, _ll, _f = C.c_void_p, C.c_int, C.c_longlong, C.c_float

# name -> (restype, argtypes); mirrors include/fosvos_b200.h one to one
SIGNATURES = {
    "fosvos_abi_version": (_i, []),
    "fosvos_last_error": (C.c_char_p, []),
    "fosvos_device_check": (_i, [_i]),
    "fosvos_num_sms": (_i, [_i]),
    "fosvos_nchw_to_nhwc": (_i, [_vp, _vp, _i, _i, _i, _i, _i, _i, _vp]),
    "fosvos_nhwc_to_nchw": (_i, [_vp, _vp, _i, _i, _i, _i, _i, _i, _vp]),
    "fosvos_packed_weight_elems": (_ll, [_i, _i, _i]),
    "fosvos_pack_conv3x3_weight": (_i, [_vp, _vp, _i, _i, _i, _i, _i, _i, _vp]),
    "fosvos_pad_bias": (_i, [_vp, _vp, _i, _i, _vp]),
    "fosvos_conv3x3_simt": (_i, [_vp, _vp, _vp, _vp, _vp, _i, _i, _i, _i, _i, _i, _i, _vp]),
    "fosvos_conv3x3_tc": (_i, [_vp, _vp, _vp, _vp, _vp, _i, _i, _i, _i, _i, _i, _vp]),
    "fosvos_conv3x3_tc_pool": (_i, [_vp, _vp, _vp, _vp, _vp, _i, _i, _i, _i, _i, _i, _vp]),
    "fosvos_conv3x3_tc_pool_arg": (_i, [_vp, _vp, _vp, _vp, _vp, _vp, _i, _i, _i, _i, _i, _i, _vp]),
    "fosvos_conv3x3_side_tc_supported": (_i, [_i]),
    "fosvos_conv3x3_side_tc": (_i, [_vp, _vp, _vp, _vp, _vp, _vp, _i, _i, _i, _i, _vp]),
    "fosvos_split_pairs": (_i, [_i]),
    "fosvos_split_weight_term": (_i, [_i, _i]),
    "fosvos_split_nchw": (_i, [_vp, _vp, _i, _i, _i, _i, _i, _i, _vp]),
    "fosvos_packed_weight_split_elems": (_ll, [_i, _i, _i]),
    "fosvos_pack_conv3x3_weight_split": (_i, [_vp, _vp, _i, _i, _i, _i, _i, _vp]),
    "fosvos_conv3x3_tc_split": (_i, [_vp, _vp, _vp, _vp, _vp, _i, _i, _i, _i, _i, _i, _i, _vp]),
    "fosvos_maxpool2x2_split": (_i, [_vp, _vp, _i, _i, _i, _i, _i, _vp]),
    "fosvos_conv3x3_wgrad_simt": (_i, [_vp, _vp, _vp, _vp, _i, _i, _i, _i, _i, _i, _i, _i, _vp]),
    "fosvos_conv3x3_wgrad_tc_workspace_bytes": (C.c_size_t, [_i, _i]),
    "fosvos_conv3x3_wgrad_tc": (_i, [_vp, _vp, _vp, _vp, _vp, _i, _i, _i, _i, _i, _i, _i, _vp]),
    "fosvos_conv3x3_wgrad_tc_accumulate": (_i, [_vp, _vp, _vp, _vp, _i, _i, _i, _i, _i, _i, _vp]),
    "fosvos_conv3x3_wgrad_tc_finish": (_i, [_vp, _vp, _i, _i, _i, _i, _i, _vp]),
    "fosvos_conv3x3_wgrad_tc_orientation": (_i, [_i, _i]),
    "fosvos_fold_tile_count": (_i, [_i, _i]),
    "fosvos_repack_tile_count": (_i, [_i, _i]),
    "fosvos_conv_step_all": (_i, [_vp, _i, _vp, _i, _f, _vp]),
    "fosvos_wgrad_fold_all": (_i, [_vp, _i, _vp, _i, _vp]),
    "fosvos_repack_all": (_i, [_vp, _i, _vp, _i, _vp]),
    "fosvos_adam_chunk_elems": (_i, []),
    "fosvos_adam_step": (_i, [_vp, _i, _vp, _i, C.c_float, C.c_float, C.c_float, _vp, _i, _vp]),
    "fosvos_pixel_loss": (_i, [_vp, _vp, C.c_longlong, _i, _i, C.c_float, _vp, _vp, _vp]),
    "fosvos_relu_fwd": (_i, [_vp, _vp, _ll, _vp]),
    "fosvos_relu_bwd": (_i, [_vp, _vp, _vp, _ll, _vp]),
    "fosvos_taylor_rank": (_i, [_vp, _vp, _vp, _i, _i, _i, _i, _i, _i, _vp]),
    "fosvos_ingest_u8": (_i, [_vp, _vp, _i, _i, _i, C.POINTER(C.c_float), _i, _vp]),
    "fosvos_maxpool2x2_fwd": (_i, [_vp, _vp, _i, _i, _i, _i, _i, _vp]),
    "fosvos_maxpool2x2_bwd": (_i, [_vp, _vp, _vp, _i, _i, _i, _i, _i, _vp]),
    "fosvos_maxpool2x2_bwd_add": (_i, [_vp, _vp, _vp, _vp, _i, _i, _i, _i, _i, _vp]),
    "fosvos_maxpool2x2_bwd_arg": (_i, [_vp, _vp, _vp, _vp, _i, _i, _i, _i, _i, _vp]),
    "fosvos_side_params_bytes": (C.c_size_t, []),
    "fosvos_side_workspace_bytes": (C.c_size_t, [_vp, _vp, _i]),
    "fosvos_side_upsample_plan": (_i, [_i, _i, _i, _i, _vp, _vp, _vp]),
    "fosvos_side_prepare": (_i, [_vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp]),
    "fosvos_side_check_diagonal": (_i, [_vp, _vp, _vp]),
    "fosvos_side_params_separable_flag": (_i, []),
    "fosvos_side_fwd": (_i, [_vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _i, _i, _i, _i, _i, _vp]),
    "fosvos_side_params_heads_offset": (_i, [_i]),
    "fosvos_side_fwd_heads_done": (_i, [_vp, _vp, _vp, _vp, _vp, _vp, _vp, _i, _i, _i, _i, _vp]),
    "fosvos_side_bwd": (_i, [_vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _i, _i, _i, _i, _vp]),
    "fosvos_bal_loss_stats_bytes": (C.c_size_t, []),
    "fosvos_bal_loss_fwd": (_i, [_vp, _vp, _ll, _i, _vp, _vp, _vp]),
    "fosvos_bal_loss_fwd_bwd": (_i, [_vp, _vp, _ll, _i, _vp, _vp, _vp, _f, _vp, _vp]),
    "fosvos_bal_loss_fwd_bwd_frames": (_i, [_vp, _vp, _ll, _i, _i, _vp, _ll, _vp, _vp, _f, _vp, _vp]),
    "fosvos_bal_loss_bwd": (_i, [_vp, _vp, _ll, _i, _vp, _vp, _f, _vp, _vp]),
    "fosvos_loss_accumulate": (_i, [_vp, _vp, _vp, _i, _i, _vp]),
    "fosvos_loss_window_finish": (_i, [_vp, _i, _vp, _vp, _vp]),
    "fosvos_sgd_chunk_elems": (_i, []),
    "fosvos_sgd_step": (_i, [_vp, _i, _vp, _i, _f, _i, _vp]),
    "fosvos_mask_iou": (_i, [_vp, _vp, _ll, _i, _vp, _vp]),
    "fosvos_sigmoid_threshold": (_i, [_vp, _vp, _vp, _ll, _vp]),
}

_lib: Optional[C.CDLL] = None


def lib() -> C.CDLL:
    """The loaded library.  Raises (never falls back) when it is missing."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise RuntimeError(
                f"fosvos_b200: {LIB_PATH} is missing. Build it with `python -m fosvos_b200.build` "
                "(nvcc, sm_100a). There is no CPU or PyTorch fallback for this path.")
        l = C.CDLL(LIB_PATH)
        for name, (res, args) in SIGNATURES.items():
            fn = getattr(l, name)           # AttributeError if the symbol is not exported
            fn.restype = res
            fn.argtypes = args
        _lib = l
    return _lib


def last_error() -> str:
    return lib().fosvos_last_error().decode("utf-8", "replace")


# number of ABI compute calls issued (each enqueues >= 1 kernel); graph replays add their node count
CALLS = [0]


def check(rc: int, what: str = "") -> None:
    CALLS[0] += 1
    if rc != 0:
        raise RuntimeError(f"fosvos_b200 {what} failed (status {rc}): {last_error()}")


def require_device(device: torch.device) -> None:
    """Fail loudly unless `device` is a CUDA sm_100 device."""
    if device.type != "cuda":
        raise RuntimeError(f"fosvos_b200 runs on CUDA sm_100 devices only, got tensor on '{device}'. "
                           "There is no CPU fallback; move the module and inputs to a B200.")
    check(lib().fosvos_device_check(device.index if device.index is not None else torch.cuda.current_device()),
          "device_check")


def stream() -> int:
    """The current stream of the current device (the op wrappers make the tensors' device current first)."""
    return torch.cuda.current_stream().cuda_stream


def _first_cuda_device(args, kwargs):
    for a in list(args) + list(kwargs.values()):
        if isinstance(a, torch.Tensor):
            if a.is_cuda:
                return a.device
        elif isinstance(a, (list, tuple)):
            for b in a:
                if isinstance(b, torch.Tensor) and b.is_cuda:
                    return b.device
    return None


def on_tensor_device(fn):
    """Decorator for the op wrappers: run `fn` with the device of its first CUDA tensor argument made current, so
    that the launch stream, the tensor maps and the per-device function attributes all belong to the GPU that owns
    the data (a process may drive several GPUs, and gloo-backend ranks never call set_device)."""
    import functools

    @functools.wraps(fn)
    def wrapper(*args, **kwargs):
        dev = _first_cuda_device(args, kwargs)
        if dev is None or dev.index is None or dev.index == torch.cuda.current_device():
            return fn(*args, **kwargs)
        with torch.cuda.device(dev):
            return fn(*args, **kwargs)
    return wrapper


def ptr(t: Optional[torch.Tensor]) -> Optional[int]:
    return None if t is None else t.data_ptr()


def ptr_array(ts: Sequence[Optional[torch.Tensor]]):
    """Host array of device pointers (for the `const T* const*` arguments)."""
    arr = (C.c_void_p * len(ts))()
    for i, t in enumerate(ts):
        arr[i] = None if t is None else t.data_ptr()
    return arr


def int_array(vals: Sequence[int]):
    return (C.c_int * len(vals))(*vals)


def dtype_code(dt: torch.dtype) -> int:
    if dt == torch.float32:
        return F32
    if dt == torch.bfloat16:
        return BF16
    raise RuntimeError(f"fosvos_b200: unsupported activation dtype {dt}")
