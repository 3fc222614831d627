"""ctypes binding of ``libi2l_b200.so`` (C-ABI declared in ``include/i2l_b200.h``).

The library is built in-tree by ``hmer-img2latex_b200/csrc/Makefile`` with nvcc for
sm_100a.  There is NO fallback: if the shared object is missing or an entry point
fails, a ``RuntimeError`` is raised (the message comes from ``i2l_last_error``).
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess
import threading

_HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(_HERE, "csrc")
LIB_PATH = os.environ.get("I2L_LIB", os.path.join(CSRC, "libi2l_b200.so"))

MAX_CONV = 8
MAX_LSTM = 8
MAX_RESNET_CONVS = 160

FP32, BF16, BF16_STREAMED = 0, 1, 2
STOP_NONE, STOP_ALL_END_SAME_STEP, STOP_ALL_FINISHED_STICKY = 0, 1, 2
PRECISIONS = {"fp32": FP32, "bf16": BF16}
IN_F32, IN_BF16, IN_U8 = 0, 1, 2
NORM_PM1, NORM_MEANSTD = 0, 1
FILTER_LANCZOS, FILTER_BICUBIC = 0, 1
RESIZE_ASPECT_PAD_CROP, RESIZE_STRETCH = 0, 1

_fp = C.c_void_p  # all device pointers travel as void*


class CnnDesc(C.Structure):
    _fields_ = [("img_height", C.c_int32), ("img_width", C.c_int32), ("channels", C.c_int32),
                ("n_conv", C.c_int32), ("filters", C.c_int32 * MAX_CONV), ("kernel_size", C.c_int32),
                ("pool_size", C.c_int32), ("embedding_dim", C.c_int32), ("precision", C.c_int32)]


class CnnParams(C.Structure):
    _fields_ = [("conv_w", _fp * MAX_CONV), ("conv_b", _fp * MAX_CONV), ("fc_w", _fp), ("fc_b", _fp)]


class ResnetDesc(C.Structure):
    _fields_ = [("depth", C.c_int32), ("img_height", C.c_int32), ("embedding_dim", C.c_int32),
                ("precision", C.c_int32)]


class ResnetParams(C.Structure):
    _fields_ = [("n_convs", C.c_int32), ("conv_w", _fp * MAX_RESNET_CONVS), ("bn_weight", _fp * MAX_RESNET_CONVS),
                ("bn_bias", _fp * MAX_RESNET_CONVS), ("bn_mean", _fp * MAX_RESNET_CONVS),
                ("bn_var", _fp * MAX_RESNET_CONVS), ("fc_w", _fp), ("fc_b", _fp)]


class ImageDesc(C.Structure):
    _fields_ = [("src_offset", C.c_int64), ("height", C.c_int32), ("width", C.c_int32)]


class DecDesc(C.Structure):
    _fields_ = [("vocab_size", C.c_int32), ("embedding_dim", C.c_int32), ("hidden_dim", C.c_int32),
                ("lstm_layers", C.c_int32), ("attention", C.c_int32), ("precision", C.c_int32)]


class DecParams(C.Structure):
    _fields_ = [("embedding", _fp), ("w_ih", _fp * MAX_LSTM), ("w_hh", _fp * MAX_LSTM), ("b_ih", _fp * MAX_LSTM),
                ("b_hh", _fp * MAX_LSTM), ("out_w", _fp), ("out_b", _fp)]


# name -> (restype, argtypes); mirrors include/i2l_b200.h one to one.
SIGNATURES = {
    "i2l_version": (C.c_char_p, []),
    "i2l_last_error": (C.c_char_p, []),
    "i2l_device_check": (C.c_int, []),
    "i2l_cnn_tensor_core_path": (C.c_int, [C.POINTER(CnnDesc)]),
    "i2l_cnn_packed_bytes": (C.c_size_t, [C.POINTER(CnnDesc)]),
    "i2l_cnn_pack": (C.c_int, [C.POINTER(CnnDesc), C.POINTER(CnnParams), _fp, C.c_size_t, _fp]),
    "i2l_cnn_workspace_bytes": (C.c_size_t, [C.POINTER(CnnDesc), C.c_int32]),
    "i2l_cnn_encoder_fwd": (C.c_int, [C.POINTER(CnnDesc), _fp, _fp, C.c_int32, _fp, _fp, C.c_size_t, _fp]),
    "i2l_cnn_encoder_fwd_in": (C.c_int, [C.POINTER(CnnDesc), _fp, _fp, C.c_int32, C.c_int32, _fp, _fp, C.c_size_t,
                                         _fp]),
    "i2l_cnn_encoder_fwd_u8": (C.c_int, [C.POINTER(CnnDesc), _fp, _fp, C.c_int32, C.POINTER(C.c_float), C.POINTER(C.c_float),
                                         C.c_int32, _fp, _fp, C.c_size_t, _fp]),
    "i2l_normalize_u8": (C.c_int, [_fp, C.c_int32, C.c_int32, C.c_int32, C.c_int32, C.c_int32, C.c_int32,
                                   C.POINTER(C.c_float), C.POINTER(C.c_float), _fp, C.c_int32, _fp]),
    "i2l_resize_plan_bytes": (C.c_size_t, [C.POINTER(ImageDesc), C.c_int32, C.c_int32, C.c_int32, C.c_int32, C.c_int32,
                                           C.c_int32, C.c_int32]),
    "i2l_resize_plan_build": (C.c_int, [C.POINTER(ImageDesc), C.c_int32, C.c_int32, C.c_int32, C.c_int32, C.c_int32,
                                        C.c_int32, C.c_int32, _fp, C.c_size_t]),
    "i2l_pack_images": (C.c_int, [C.POINTER(C.c_void_p), C.POINTER(ImageDesc), C.c_int32, C.c_int32, _fp]),
    "i2l_resize_workspace_bytes": (C.c_size_t, [_fp]),
    "i2l_resize_pad_u8": (C.c_int, [_fp, _fp, _fp, _fp, _fp, C.c_size_t, _fp]),
    "i2l_resnet_num_convs": (C.c_int32, [C.c_int32]),
    "i2l_resnet_packed_bytes": (C.c_size_t, [C.POINTER(ResnetDesc)]),
    "i2l_resnet_pack": (C.c_int, [C.POINTER(ResnetDesc), C.POINTER(ResnetParams), _fp, C.c_size_t, _fp]),
    "i2l_resnet_workspace_bytes": (C.c_size_t, [C.POINTER(ResnetDesc), C.c_int32, C.c_int32]),
    "i2l_resnet_encoder_fwd": (C.c_int, [C.POINTER(ResnetDesc), _fp, _fp, C.c_int32, C.c_int32, _fp, _fp,
                                         C.c_size_t, _fp]),
    "i2l_attention_workspace_bytes": (C.c_size_t, [C.c_int32, C.c_int32, C.c_int32, C.c_int32]),
    "i2l_attention_fwd": (C.c_int, [C.c_int32, C.c_int32, _fp, _fp, _fp, _fp, _fp, C.c_int32, C.c_int32, _fp, _fp,
                                    C.c_size_t, _fp]),
    "i2l_dec_packed_bytes": (C.c_size_t, [C.POINTER(DecDesc)]),
    "i2l_dec_pack": (C.c_int, [C.POINTER(DecDesc), C.POINTER(DecParams), _fp, C.c_size_t, _fp]),
    "i2l_dec_workspace_bytes": (C.c_size_t, [C.POINTER(DecDesc), C.c_int32, C.c_int32]),
    "i2l_decode_step": (C.c_int, [C.POINTER(DecDesc), _fp, _fp, _fp, C.c_int32, _fp, _fp, _fp, _fp, _fp, _fp, _fp,
                                  C.c_size_t, _fp]),
    "i2l_dec_forward_workspace_bytes": (C.c_size_t, [C.POINTER(DecDesc), C.c_int32, C.c_int32]),
    "i2l_decoder_forward": (C.c_int, [C.POINTER(DecDesc), _fp, _fp, _fp, C.c_int32, C.c_int32, _fp, _fp, _fp, _fp, _fp,
                                      _fp, _fp, C.c_size_t, _fp]),
    "i2l_decode_greedy": (C.c_int, [C.POINTER(DecDesc), _fp, _fp, C.c_int32, C.c_int32, C.c_int32, C.c_int32,
                                    C.c_float, C.c_int32, _fp, _fp, _fp, _fp, C.c_size_t, _fp]),
    "i2l_decode_sample": (C.c_int, [C.POINTER(DecDesc), _fp, _fp, C.c_int32, C.c_int32, C.c_int32, C.c_int32,
                                    C.c_float, C.c_int32, C.c_float, C.c_uint64, C.c_uint64, _fp, _fp, _fp, _fp,
                                    _fp, _fp, C.c_size_t, _fp]),
    "i2l_decode_beam": (C.c_int, [C.POINTER(DecDesc), _fp, _fp, C.c_int32, C.c_int32, C.c_int32, C.c_int32,
                                  C.c_int32, _fp, _fp, _fp, _fp, _fp, _fp, _fp, _fp, _fp, C.c_size_t, _fp]),
    "i2l_token_exchange_buffer_bytes": (C.c_size_t, [C.c_int32, C.c_int32, C.c_int32]),
    "i2l_token_exchange_write": (C.c_int, [_fp, _fp, _fp, C.c_int32, C.c_int32, C.c_int32, C.c_int32, C.c_int32,
                                           C.POINTER(C.c_void_p), C.c_uint32, _fp]),
    "i2l_token_exchange_read": (C.c_int, [_fp, C.c_int32, C.c_int32, C.c_int32, C.c_uint32, _fp, _fp, _fp, _fp, _fp]),
    "i2l_sequence_metrics": (C.c_int, [_fp, C.c_int32, _fp, _fp, C.c_int32, _fp, C.c_int32, C.c_int32, _fp, _fp]),
    "i2l_xent_workspace_bytes": (C.c_size_t, [C.c_int32]),
    "i2l_xent_metrics": (C.c_int, [_fp, _fp, C.c_int32, C.c_int32, C.c_int64, C.c_float, _fp, _fp, _fp, C.c_size_t, _fp]),
    "i2l_filter_ids": (C.c_int, [_fp, C.c_int32, _fp, C.c_int32, C.POINTER(C.c_int64), C.c_int32, _fp, C.c_int32, _fp, _fp]),
    "i2l_launch_count": (C.c_longlong, []),
    "i2l_prof_enable": (None, [C.c_int]),
    "i2l_prof_reset": (None, []),
    "i2l_prof_count": (C.c_int, []),
    "i2l_prof_get": (C.c_int, [C.c_int, C.c_char_p, C.c_int, C.POINTER(C.c_int), C.POINTER(C.c_float)]),
}

_lib = None
_lock = threading.Lock()


def build(verbose: bool = False) -> str:
    """Compile the CUDA sources for sm_100a (nvcc cross-compiles without a GPU)."""
    r = subprocess.run(["make", "-C", CSRC, "-j8"], capture_output=True, text=True)
    if verbose or r.returncode != 0:
        print(r.stdout[-4000:])
        print(r.stderr[-4000:])
    if r.returncode != 0:
        raise RuntimeError("building libi2l_b200.so failed (see output above)")
    return LIB_PATH


def lib() -> C.CDLL:
    """Load the native library; raises loudly when it is missing (no fallback path)."""
    global _lib
    if _lib is not None:
        return _lib
    with _lock:
        if _lib is None:
            if not os.path.exists(LIB_PATH):
                raise RuntimeError(
                    f"{LIB_PATH} not found: the sm_100a CUDA library is required and there is no CPU / "
                    f"PyTorch fallback.  Build it with `python -c 'import __graft_entry__ as g; g.build()'` "
                    f"or `make -C {CSRC}`.")
            l = C.CDLL(LIB_PATH)
            for name, (res, args) in SIGNATURES.items():
                fn = getattr(l, name)  # AttributeError if the symbol is not exported
                fn.restype = res
                fn.argtypes = args
            _lib = l
    return _lib


def last_error() -> str:
    return lib().i2l_last_error().decode("utf-8", "replace")


def check(status: int, what: str) -> None:
    if status != 0:
        raise RuntimeError(f"{what} failed with status {status}: {last_error()}")


def ptr(t) -> C.c_void_p:
    """Device pointer of a torch tensor (or NULL for None)."""
    if t is None:
        return C.c_void_p(0)
    return C.c_void_p(t.data_ptr())


def stream_ptr(device=None) -> C.c_void_p:
    import torch
    return C.c_void_p(torch.cuda.current_stream(device).cuda_stream)


def prof_results() -> dict:
    """{kernel name: (launches, total_ms)} from the library's CUDA-event kernel timers."""
    l = lib()
    out = {}
    for i in range(l.i2l_prof_count()):
        name = C.create_string_buffer(64)
        cnt, ms = C.c_int(0), C.c_float(0.0)
        check(l.i2l_prof_get(i, name, 64, C.byref(cnt), C.byref(ms)), "i2l_prof_get")
        if cnt.value:
            out[name.value.decode()] = (cnt.value, ms.value)
    return out
