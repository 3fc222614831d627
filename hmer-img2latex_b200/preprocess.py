"""Device-side image preparation: the geometry + pixel arithmetic of the reference's
``load_image`` (img2latex/data/utils.py:18-90), ``ResizeWithAspectRatio``
(img2latex/data/transforms.py:9-56) and the PIL branch of ``Predictor._prepare_image``
(img2latex/training/predictor.py:427-446) for a ragged BATCH of uint8 images.

Host work is limited to what cannot be done on the GPU bit-exactly or at all: decoding files,
packing the raw pixels into one pinned buffer and building the filter-weight plan
(``i2l_resize_plan_build``: Pillow's double-precision weight computation).  Pixels + plan cross
PCIe in ONE copy; ``i2l_resize_pad_u8`` then resamples, pads / crops and writes planar uint8
``(B,C,H,W)``, which ``i2l_normalize_u8`` / ``i2l_cnn_encoder_fwd_u8`` consume.  There is no CPU
fallback: without the CUDA library every call raises.
"""
from __future__ import annotations

import ctypes as C
import threading
from typing import List, Optional, Sequence, Tuple, Union

import numpy as np
import torch

from . import _native as N
from .model.encoder import normalize_u8

_FILTERS = {"lanczos": N.FILTER_LANCZOS, "bicubic": N.FILTER_BICUBIC}
_MODES = {"aspect": N.RESIZE_ASPECT_PAD_CROP, "stretch": N.RESIZE_STRETCH}


def _as_u8_array(img) -> np.ndarray:
    """PIL image / ndarray -> (H,W) or (H,W,3) uint8 array in PIL's memory order."""
    if isinstance(img, np.ndarray):
        a = img
    elif isinstance(img, torch.Tensor):
        a = img.detach().cpu().numpy()
    else:                                        # PIL.Image.Image (duck-typed: no hard PIL dependency)
        if img.mode not in ("L", "RGB"):
            img = img.convert("RGB")
        a = np.asarray(img)
    if a.dtype != np.uint8:
        raise TypeError(f"image pixels must be uint8, got {a.dtype}")
    if a.ndim == 3 and a.shape[2] == 1:
        a = a[:, :, 0]
    if a.ndim not in (2, 3) or (a.ndim == 3 and a.shape[2] != 3):
        raise ValueError(f"images must be (H,W) or (H,W,3) uint8 arrays, got shape {a.shape}")
    return np.ascontiguousarray(a)


class _PlanHeader(C.Structure):       # mirrors PlanHeader in csrc/preprocess.cu (introspection for the tests)
    _fields_ = [("magic", C.c_uint32), ("n", C.c_int32), ("src_channels", C.c_int32), ("out_channels", C.c_int32),
                ("to_gray", C.c_int32), ("target_h", C.c_int32), ("target_w", C.c_int32), ("filter", C.c_int32),
                ("mode", C.c_int32), ("max_inter_pixels", C.c_int32), ("plan_bytes", C.c_uint64),
                ("workspace_bytes", C.c_uint64), ("images_off", C.c_uint64)]


class _ImagePlan(C.Structure):        # mirrors ImagePlan in csrc/preprocess.cu
    _fields_ = [("src_offset", C.c_int64), ("inter_off", C.c_uint64), ("kh_off", C.c_uint64), ("bh_off", C.c_uint64),
                ("kv_off", C.c_uint64), ("bv_off", C.c_uint64), ("h", C.c_int32), ("w", C.c_int32),
                ("new_w", C.c_int32), ("new_h", C.c_int32), ("left", C.c_int32), ("kept_w", C.c_int32),
                ("kept_h", C.c_int32), ("need_h", C.c_int32), ("need_v", C.c_int32), ("pitch", C.c_int32),
                ("pad_", C.c_int32)]


class _PinnedPool:
    """Two reusable pinned staging buffers (cudaHostAlloc of tens of MB per batch would cost more than the
    copy).  A slot belongs to the plan that took it until that plan's `run()` has queued its H2D copy (`mark`); it is
    handed out again only after that copy has completed.  A plan created while both slots are still held by un-run
    plans (or by other threads: loader threads, two Predictors) gets a private, unpooled pinned buffer instead of
    overwriting pixels somebody still needs.  All state changes happen under one lock."""

    def __init__(self):
        import threading
        self._lock = threading.Lock()
        self._bufs = [None, None]
        self._events = [None, None]
        self._held = [False, False]
        self._turn = 0

    def get(self, nbytes: int):
        with self._lock:
            order = (self._turn, self._turn ^ 1)
            i = next((k for k in order if not self._held[k]), -1)
            if i >= 0:
                self._held[i] = True
                self._turn = i ^ 1
                ev, self._events[i] = self._events[i], None
        if i < 0:                                                 # both slots belong to plans that have not run yet
            return -1, torch.empty(max(int(nbytes), 1), dtype=torch.uint8, pin_memory=torch.cuda.is_available())
        if ev is not None:
            ev.synchronize()                                      # the copy that last read this slot (outside the lock)
        b = self._bufs[i]
        if b is None or b.numel() < nbytes:
            b = torch.empty(max(int(nbytes * 1.25), 1 << 20), dtype=torch.uint8, pin_memory=torch.cuda.is_available())
            self._bufs[i] = b                                     # only the holder of slot i touches _bufs[i]
        return i, b[:nbytes]

    def mark(self, i: int, event) -> None:
        """the plan's H2D copy has been queued: the slot is free for the next taker once `event` has fired"""
        if i < 0:
            return
        with self._lock:
            self._events[i] = event
            self._held[i] = False

    def release(self, i: int) -> None:
        """a plan that is dropped without having run gives its slot back"""
        if i < 0:
            return
        with self._lock:
            self._held[i] = False


_POOL = _PinnedPool()


class ResizePlan:
    """Packed pixels + resize plan of one batch in a single pinned host buffer (a slice of a two-slot
    staging pool: `run` a plan before creating two further ones)."""

    def __init__(self, images: Sequence, target_height: int, target_width: int, to_gray: bool = False,
                 resample: str = "lanczos", mode: str = "aspect"):
        arrs = [_as_u8_array(im) for im in images]
        chans = {1 if a.ndim == 2 else 3 for a in arrs}
        if len(chans) > 1:
            raise ValueError("one batch must hold images of one mode (all L or all RGB); convert or split the batch")
        self.n = len(arrs)
        self.src_channels = chans.pop() if chans else 1
        self.to_gray = bool(to_gray) and self.src_channels == 3
        self.out_channels = 1 if self.to_gray else self.src_channels
        self.target_height, self.target_width = int(target_height), int(target_width)
        lib = N.lib()
        descs = (N.ImageDesc * max(self.n, 1))()
        off = 0
        for i, a in enumerate(arrs):
            descs[i].src_offset, descs[i].height, descs[i].width = off, a.shape[0], a.shape[1]
            off += (a.size + 15) // 16 * 16
        self.pixel_bytes = off
        args = (descs, self.n, self.src_channels, int(self.to_gray), self.target_height, self.target_width,
                _FILTERS[resample], _MODES[mode])
        self.plan_bytes = lib.i2l_resize_plan_bytes(*args)
        if self.plan_bytes == 0:
            raise ValueError(f"resize plan rejected: {N.last_error()}")     # Pillow raises ValueError for empty targets
        self._slot, self.host = _POOL.get(self.pixel_bytes + self.plan_bytes)
        self.plan_ptr = self.host.data_ptr() + self.pixel_bytes
        # the weight tables are built by the library's own threads (ctypes releases the GIL for the call) while this
        # thread packs the pixels into the other half of the staging buffer
        status = []

        def build():                                     # the error text is thread-local: read it on this thread
            rc = lib.i2l_resize_plan_build(*args, C.c_void_p(self.plan_ptr), self.plan_bytes)
            status.append((rc, N.last_error() if rc else ""))

        builder = threading.Thread(target=build)
        builder.start()
        if self.n:
            ptrs = (C.c_void_p * self.n)(*[a.ctypes.data for a in arrs])
            N.check(lib.i2l_pack_images(ptrs, descs, self.n, self.src_channels, C.c_void_p(self.host.data_ptr())),
                    "i2l_pack_images")
        builder.join()
        if not status or status[0][0] != 0:
            raise RuntimeError(f"i2l_resize_plan_build failed: {status[0][1] if status else 'builder thread died'}")
        self.workspace_bytes = lib.i2l_resize_workspace_bytes(C.c_void_p(self.plan_ptr))

    def describe(self) -> List[dict]:
        """Per-image geometry and filter tables read back from the plan blob (host only; used by the tests):
        bounds (out,2) int32 and weights (out, ksize) int32 per pass, or None for a skipped pass."""
        blob = self.host.numpy()[self.pixel_bytes:]
        hd = _PlanHeader.from_buffer_copy(blob[:C.sizeof(_PlanHeader)].tobytes())
        out = []
        for i in range(hd.n):
            o = hd.images_off + i * C.sizeof(_ImagePlan)
            ip = _ImagePlan.from_buffer_copy(blob[o:o + C.sizeof(_ImagePlan)].tobytes())
            d = {k: getattr(ip, k) for k, _ in _ImagePlan._fields_}
            for tag, need, k_off, b_off, n_in, n_out in (("h", ip.need_h, ip.kh_off, ip.bh_off, ip.w, ip.new_w),
                                                        ("v", ip.need_v, ip.kv_off, ip.bv_off, ip.h, ip.new_h)):
                if not need:
                    d[f"bounds_{tag}"] = d[f"weights_{tag}"] = None
                    continue
                bnd = blob[b_off:b_off + 8 * n_out].view(np.int32).reshape(n_out, 2).copy()
                ks = (int(b_off) - int(k_off)) // (4 * n_out)              # rows of the transposed table (16-byte padded)
                kk = blob[k_off:k_off + 4 * ks * n_out].view(np.int32).reshape(ks, n_out).T.copy()
                d[f"bounds_{tag}"], d[f"weights_{tag}"] = bnd, kk
            out.append(d)
        return out

    def __del__(self):
        try:
            _POOL.release(getattr(self, "_slot", -1))
        except Exception:
            pass

    def run(self, device, workspace: Optional[torch.Tensor] = None, out: Optional[torch.Tensor] = None) -> torch.Tensor:
        """One H2D copy (pixels + plan) and two launches -> uint8 (n, C_out, Ht, Wt) on `device`."""
        device = torch.device(device)
        if device.type != "cuda":
            raise RuntimeError("image preparation needs a CUDA (sm_100a) device; there is no CPU fallback")
        with torch.cuda.device(device):
            dev = self.host.to(device, non_blocking=True)
            ev = torch.cuda.Event()
            ev.record(torch.cuda.current_stream(device))
            _POOL.mark(self._slot, ev)
            self._slot = -1                                       # a second run() re-reads the host slice it no longer owns:
            # allowed while no other plan has taken the slot since (as before); the pool no longer tracks this plan
            if out is None:
                out = torch.empty(self.n, self.out_channels, self.target_height, self.target_width, dtype=torch.uint8,
                                  device=device)
            if workspace is None or workspace.numel() < self.workspace_bytes:
                workspace = torch.empty(max(self.workspace_bytes, 256), dtype=torch.uint8, device=device)
            N.check(N.lib().i2l_resize_pad_u8(N.ptr(dev), C.c_void_p(self.plan_ptr),
                                              C.c_void_p(dev.data_ptr() + self.pixel_bytes), N.ptr(out), N.ptr(workspace),
                                              workspace.numel(), N.stream_ptr(device)), "i2l_resize_pad_u8")
            dev.record_stream(torch.cuda.current_stream(device))
            workspace.record_stream(torch.cuda.current_stream(device))
        return out


class ResizeWithAspectRatio:
    """Batched, device-side ``ResizeWithAspectRatio`` (data/transforms.py:9-56): resize each image to the
    target height keeping its aspect ratio (LANCZOS), then pad right with white or centre-crop to the
    target width.  ``__call__`` takes a sequence of PIL images / uint8 arrays and returns a uint8 CUDA
    tensor ``(B, C, target_height, target_width)`` whose pixels equal the PIL result bit for bit."""

    def __init__(self, target_height: int, target_width: int, device=None):
        self.target_height, self.target_width = target_height, target_width
        self.device = device

    def __call__(self, images: Sequence, to_gray: bool = False) -> torch.Tensor:
        dev = self.device if self.device is not None else torch.device("cuda", torch.cuda.current_device())
        return ResizePlan(images, self.target_height, self.target_width, to_gray=to_gray).run(dev)


def open_image(path: str, channels: int) -> np.ndarray:
    """File decoding (host I/O; data/utils.py:37-44): PIL open + mode conversion, as uint8 array."""
    from PIL import Image                          # host-side decoder only
    img = Image.open(path)
    if channels == 1 and img.mode != "L":
        img = img.convert("L")
    elif channels == 3 and img.mode != "RGB":
        img = img.convert("RGB")
    return np.asarray(img)


def load_images(images: Sequence[Union[str, np.ndarray]], img_size: Tuple[int, int] = (64, 800), channels: int = 1,
                normalize: bool = True, device=None, return_uint8: bool = False) -> torch.Tensor:
    """Batched ``load_image`` (data/utils.py:18-90): paths or already-decoded uint8 arrays ->
    ``(B, channels, H, W)`` float32 on the device ([0,1] scaling, then [-1,1] for 1 channel or ImageNet
    mean / std for RGB when ``normalize``).  RGB arrays are converted to L on the device when
    ``channels == 1`` (Pillow's ``convert('L')`` arithmetic); L arrays are replicated for ``channels == 3``
    (``convert('RGB')``)."""
    dev = torch.device(device) if device is not None else torch.device("cuda", torch.cuda.current_device())
    arrs: List[np.ndarray] = [open_image(im, channels) if isinstance(im, str) else _as_u8_array(im) for im in images]
    gray_src = [a.ndim == 2 for a in arrs]
    if channels == 3 and any(gray_src):            # convert("RGB") of an L image replicates the channel
        arrs = [np.repeat(a[:, :, None], 3, axis=2) if g else a for a, g in zip(arrs, gray_src)]
    if channels == 1 and any(gray_src) and not all(gray_src):
        raise ValueError("mixed L / RGB batch for channels=1: decode with open_image(path, 1) or split the batch")
    u8 = ResizePlan(arrs, img_size[0], img_size[1], to_gray=(channels == 1)).run(dev)
    if return_uint8:
        return u8
    if not normalize:                              # data/utils.py:68-69 only: x / 255
        return normalize_u8(u8, "meanstd", mean=(0.0,) * 4, std=(1.0,) * 4)      # (x/255 - 0) / 1 == x/255 bit for bit
    return normalize_u8(u8, "pm1" if channels == 1 else "meanstd")
