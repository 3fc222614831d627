"""Drop-in encoders: same constructor signatures, attributes and ``state_dict`` keys as
the reference's ``img2latex/model/encoder.py`` (CNNEncoder 16-129, ResNetEncoder
132-249); ``forward`` runs the sm_100a kernels through the C-ABI
(``i2l_cnn_encoder_fwd`` / ``i2l_resnet_encoder_fwd``).  The torch sub-modules exist
only to own the parameters under the reference's names -- they are never called.
"""
from __future__ import annotations

import ctypes as C
from typing import List, Optional

import torch
import torch.nn as nn

from .. import _native as N
from .. import ops as _ops  # noqa: F401  (registers torch.ops.i2l.*)
from ._common import Workspace, default_precision, f32c, params_key, require_cuda

_RESNET_DEPTH = {"resnet18": 18, "resnet34": 34, "resnet50": 50, "resnet101": 101, "resnet152": 152}


IMAGENET_MEAN = (0.485, 0.456, 0.406)          # data/utils.py:77-78
IMAGENET_STD = (0.229, 0.224, 0.225)


def normalize_u8(pixels: torch.Tensor, mode: str = "pm1", mean=IMAGENET_MEAN, std=IMAGENET_STD,
                 channels_last: bool = False, out_dtype: torch.dtype = torch.float32) -> torch.Tensor:
    """uint8 pixels on the device -> normalised (B,C,H,W) image tensor (fp32 or bf16), the
    arithmetic of the reference's ``load_image`` / ``_prepare_image`` (data/utils.py:68-80,
    training/predictor.py:441-446) as one kernel (``i2l_normalize_u8``): ``mode="pm1"`` is
    x/255*2-1, ``mode="meanstd"`` is (x/255-mean)/std per channel."""
    require_cuda(pixels, "normalize_u8")
    if pixels.dtype != torch.uint8 or pixels.dim() != 4:
        raise RuntimeError("normalize_u8 expects a 4-D uint8 tensor")
    if mode not in ("pm1", "meanstd"):
        raise ValueError("mode must be 'pm1' or 'meanstd'")
    if out_dtype not in (torch.float32, torch.bfloat16):
        raise ValueError("out_dtype must be torch.float32 or torch.bfloat16")
    pixels = pixels.contiguous()
    if channels_last:
        B, H, W, Cc = pixels.shape
    else:
        B, Cc, H, W = pixels.shape
    if Cc > 4 or (mode == "meanstd" and (len(mean) < Cc or len(std) < Cc)):
        raise ValueError("normalize_u8 supports up to 4 channels with one mean/std per channel")
    out = torch.empty(B, Cc, H, W, dtype=out_dtype, device=pixels.device)
    m = (C.c_float * 4)(*([float(v) for v in mean[:Cc]] + [0.0] * (4 - Cc)))
    sd = (C.c_float * 4)(*([float(v) for v in std[:Cc]] + [1.0] * (4 - Cc)))
    with torch.cuda.device(pixels.device):
        N.check(N.lib().i2l_normalize_u8(N.ptr(pixels), 1 if channels_last else 0, B, Cc, H, W,
                                         N.NORM_PM1 if mode == "pm1" else N.NORM_MEANSTD, m, sd, N.ptr(out),
                                         N.IN_BF16 if out_dtype == torch.bfloat16 else N.IN_F32,
                                         N.stream_ptr(pixels.device)), "i2l_normalize_u8")
    return out


class CNNEncoder(nn.Module):
    """reference: img2latex/model/encoder.py:16-129."""

    def __init__(self, img_height: int = None, img_width: int = None, channels: int = None,
                 conv_filters: List[int] = None, kernel_size: int = None, pool_size: int = None,
                 padding: str = "same", embedding_dim: int = None, precision: Optional[str] = None):
        super().__init__()
        # defaults: encoder.py:50-64
        img_height = 64 if img_height is None else img_height
        img_width = 800 if img_width is None else img_width
        channels = 1 if channels is None else channels
        conv_filters = [32, 64, 128] if conv_filters is None else list(conv_filters)
        kernel_size = 3 if kernel_size is None else kernel_size
        pool_size = 2 if pool_size is None else pool_size
        embedding_dim = 256 if embedding_dim is None else embedding_dim
        if padding != "same":
            raise NotImplementedError("only padding='same' (the reference default, encoder.py:32) is implemented")
        if kernel_size % 2 == 0 or len(conv_filters) > N.MAX_CONV:
            raise ValueError("kernel_size must be odd and at most %d conv layers are supported" % N.MAX_CONV)
        self.img_height, self.img_width, self.channels = img_height, img_width, channels
        self.embedding_dim = embedding_dim
        self.conv_filters, self.kernel_size, self.pool_size = conv_filters, kernel_size, pool_size
        self.precision = precision or default_precision()

        layers = []
        cin, h, w = channels, img_height, img_width
        for f in conv_filters:                                   # encoder.py:78-93
            layers += [nn.Conv2d(cin, f, kernel_size, padding=kernel_size // 2), nn.ReLU(), nn.MaxPool2d(pool_size)]
            cin, h, w = f, h // pool_size, w // pool_size
        self.cnn_layers = nn.Sequential(*layers)
        self.flatten = nn.Flatten()
        self.embedding_layer = nn.Linear(cin * h * w, embedding_dim)   # encoder.py:97-106
        self.activation = nn.ReLU()
        self._packed = None
        self._packed_key = None
        self._ws = Workspace()

    # -- native plumbing -------------------------------------------------
    def _desc(self) -> N.CnnDesc:
        d = N.CnnDesc()
        d.img_height, d.img_width, d.channels = self.img_height, self.img_width, self.channels
        d.n_conv = len(self.conv_filters)
        for i, f in enumerate(self.conv_filters):
            d.filters[i] = f
        d.kernel_size, d.pool_size, d.embedding_dim = self.kernel_size, self.pool_size, self.embedding_dim
        d.precision = N.PRECISIONS[self.precision]
        return d

    def _desc_list(self) -> List[int]:
        """the descriptor as the int list the custom ops take (ops.py)."""
        return [self.img_height, self.img_width, self.channels, self.kernel_size, self.pool_size, self.embedding_dim,
                N.PRECISIONS[self.precision]] + list(self.conv_filters)

    def _weights(self):
        convs = [m for m in self.cnn_layers if isinstance(m, nn.Conv2d)]
        ts = []
        for c in convs:
            ts += [c.weight, c.bias]
        return ts + [self.embedding_layer.weight, self.embedding_layer.bias]

    def _ensure_packed(self, device):
        ts = self._weights()
        key = params_key(ts, self.precision)
        if self._packed is not None and key == self._packed_key:
            return
        lib = N.lib()
        d = self._desc()
        held = [f32c(t) for t in ts]
        p = N.CnnParams()
        n = len(self.conv_filters)
        for i in range(n):
            p.conv_w[i] = held[2 * i].data_ptr()
            p.conv_b[i] = held[2 * i + 1].data_ptr()
        p.fc_w, p.fc_b = held[2 * n].data_ptr(), held[2 * n + 1].data_ptr()
        nbytes = lib.i2l_cnn_packed_bytes(C.byref(d))
        if nbytes == 0:
            raise RuntimeError("i2l_cnn_packed_bytes rejected the configuration: " + N.last_error())
        packed = torch.empty(nbytes, dtype=torch.uint8, device=device)
        N.check(lib.i2l_cnn_pack(C.byref(d), C.byref(p), N.ptr(packed), nbytes, N.stream_ptr(device)), "i2l_cnn_pack")
        torch.cuda.current_stream(device).synchronize()   # `held` temporaries may be freed after this
        self._packed, self._packed_key = packed, key

    def forward(self, x: torch.Tensor) -> torch.Tensor:
        """(B, C, H, W) -> (B, embedding_dim); reference encoder.py:111-129."""
        require_cuda(x, "CNNEncoder.forward")
        if self.training and torch.is_grad_enabled() and any(p.requires_grad for p in self.parameters()):
            raise RuntimeError("CNNEncoder.forward: the native path builds no autograd graph (inference only); call "
                               ".eval() or wrap the call in torch.no_grad()")
        if x.dim() != 4 or tuple(x.shape[1:]) != (self.channels, self.img_height, self.img_width):
            raise RuntimeError(f"CNNEncoder expected (B,{self.channels},{self.img_height},{self.img_width}), "
                               f"got {tuple(x.shape)}")
        with torch.cuda.device(x.device):
            self._ensure_packed(x.device)
            lib = N.lib()
            d = self._desc()
            # bf16 images are consumed as they are by the tcgen05 conv1 (precision "bf16", headline
            # shape); every other dtype goes through fp32 like the reference's tensors
            if x.dtype == torch.bfloat16 and self.precision == "bf16" and self._bf16_shape():
                x = x.detach().contiguous()
            else:
                x = f32c(x)
            ws = self._ws.get(lib.i2l_cnn_workspace_bytes(C.byref(d), max(x.shape[0], 1)), x.device)
            return torch.ops.i2l.cnn_encoder_fwd(x, self._packed, ws, self._desc_list())

    def fused_u8_supported(self) -> bool:
        """True when raw uint8 pixels can be fed straight to the first convolution (``forward_u8``)."""
        return self.precision == "bf16" and self._bf16_shape()

    def forward_u8(self, pixels: torch.Tensor, normalize: str = "pm1", mean=IMAGENET_MEAN, std=IMAGENET_STD
                   ) -> torch.Tensor:
        """(B, C, H, W) uint8 pixels -> (B, embedding_dim): ``forward(normalize_u8(pixels))`` with the
        normalisation of the reference's ``load_image`` / ``_prepare_image`` fused into conv1
        (``i2l_cnn_encoder_fwd_u8``) -- the image is read once, one byte per pixel.  Shapes / precisions
        without the fused kernel run the two calls."""
        require_cuda(pixels, "CNNEncoder.forward_u8")
        if pixels.dtype != torch.uint8 or pixels.dim() != 4 or \
                tuple(pixels.shape[1:]) != (self.channels, self.img_height, self.img_width):
            raise RuntimeError(f"CNNEncoder.forward_u8 expected uint8 (B,{self.channels},{self.img_height},"
                               f"{self.img_width}), got {pixels.dtype} {tuple(pixels.shape)}")
        if normalize not in ("pm1", "meanstd"):
            raise ValueError("normalize must be 'pm1' or 'meanstd'")
        if not self.fused_u8_supported():
            return self.forward(normalize_u8(pixels, normalize, mean, std,
                                             out_dtype=torch.bfloat16 if self.precision == "bf16" else torch.float32))
        with torch.cuda.device(pixels.device):
            self._ensure_packed(pixels.device)
            lib, d = N.lib(), self._desc()
            x = pixels.contiguous()
            ws = self._ws.get(lib.i2l_cnn_workspace_bytes(C.byref(d), max(x.shape[0], 1)), x.device)
            return torch.ops.i2l.cnn_encoder_fwd_u8(x, self._packed, ws, self._desc_list(),
                                                    N.NORM_PM1 if normalize == "pm1" else N.NORM_MEANSTD,
                                                    [float(v) for v in mean[:3]], [float(v) for v in std[:3]])

    def _bf16_shape(self) -> bool:
        """the shape runs on the tcgen05 kernels in precision "bf16" (`i2l_cnn_tensor_core_path`: filters 32/64/128,
        3x3, pool 2, E = 256, 1 or 3 channels, H % 64 == 0, W % 32 == 0)"""
        d = self._desc()
        d.precision = N.BF16
        return bool(N.lib().i2l_cnn_tensor_core_path(C.byref(d)))


class ResNetEncoder(nn.Module):
    """reference: img2latex/model/encoder.py:132-249.  The torchvision trunk is built with
    ``weights=None`` unless ``pretrained=True`` (the reference always downloads ImageNet
    weights, encoder.py:185-194, which needs a network)."""

    def __init__(self, img_height: int = None, img_width: int = None, channels: int = None,
                 model_name: str = "resnet50", embedding_dim: int = None, freeze_backbone: bool = True,
                 pretrained: bool = False, precision: Optional[str] = None):
        super().__init__()
        import torchvision.models as models
        img_height = 64 if img_height is None else img_height
        img_width = 800 if img_width is None else img_width
        channels = 3 if channels is None else channels
        embedding_dim = 256 if embedding_dim is None else embedding_dim
        if model_name not in _RESNET_DEPTH:
            raise ValueError(f"Invalid ResNet model name: {model_name}")      # encoder.py:195-196
        if channels != 3:
            raise ValueError("ResNet expects 3-channel RGB images")
        self.img_height, self.img_width, self.channels = img_height, img_width, channels
        self.embedding_dim, self.model_name = embedding_dim, model_name
        self.precision = precision or default_precision()
        weights = "IMAGENET1K_V1" if pretrained else None
        backbone = getattr(models, model_name)(weights=weights)
        modules = list(backbone.children())[:-1]                              # encoder.py:198
        self.resnet = nn.Sequential(*modules)
        if freeze_backbone:                                                   # encoder.py:201-210
            for p in self.resnet.parameters():
                p.requires_grad = False
            for p in modules[-2].parameters():
                p.requires_grad = True
        self.flatten = nn.Flatten()
        feat = 512 if model_name in ("resnet18", "resnet34") else 2048        # encoder.py:219-222
        self.embedding_layer = nn.Linear(feat, embedding_dim)
        self.activation = nn.ReLU()
        self._packed = None
        self._packed_key = None
        self._ws = Workspace()

    def _desc(self) -> N.ResnetDesc:
        d = N.ResnetDesc()
        d.depth = _RESNET_DEPTH[self.model_name]
        d.img_height, d.embedding_dim = self.img_height, self.embedding_dim
        d.precision = N.PRECISIONS[self.precision]
        return d

    def _conv_bn_pairs(self):
        """(conv, bn) in the canonical order of include/i2l_b200.h: stem, then per block
        conv1, conv2[, conv3][, downsample]."""
        pairs = [(self.resnet[0], self.resnet[1])]
        for li in range(4, 8):
            for blk in self.resnet[li]:
                pairs.append((blk.conv1, blk.bn1))
                pairs.append((blk.conv2, blk.bn2))
                if hasattr(blk, "conv3"):
                    pairs.append((blk.conv3, blk.bn3))
                if blk.downsample is not None:
                    pairs.append((blk.downsample[0], blk.downsample[1]))
        return pairs

    def _ensure_packed(self, device):
        pairs = self._conv_bn_pairs()
        ts = []
        for c, b in pairs:
            ts += [c.weight, b.weight, b.bias, b.running_mean, b.running_var]
        ts += [self.embedding_layer.weight, self.embedding_layer.bias]
        key = params_key(ts, self.precision)
        if self._packed is not None and key == self._packed_key:
            return
        lib = N.lib()
        d = self._desc()
        held = [f32c(t) for t in ts]
        p = N.ResnetParams()
        p.n_convs = len(pairs)
        assert p.n_convs == lib.i2l_resnet_num_convs(d.depth)
        for i in range(len(pairs)):
            p.conv_w[i], p.bn_weight[i], p.bn_bias[i], p.bn_mean[i], p.bn_var[i] = (
                held[5 * i + j].data_ptr() for j in range(5))
        p.fc_w, p.fc_b = held[-2].data_ptr(), held[-1].data_ptr()
        nbytes = lib.i2l_resnet_packed_bytes(C.byref(d))
        packed = torch.empty(nbytes, dtype=torch.uint8, device=device)
        N.check(lib.i2l_resnet_pack(C.byref(d), C.byref(p), N.ptr(packed), nbytes, N.stream_ptr(device)),
                "i2l_resnet_pack")
        torch.cuda.current_stream(device).synchronize()
        self._packed, self._packed_key = packed, key

    def forward_buckets(self, batches: List[torch.Tensor], n_streams: int = 4, use_graphs: bool = True
                        ) -> List[torch.Tensor]:
        """Width-bucketed encoding (BASELINE configs[3]): ``[forward(x) for x in batches]`` with the buckets spread
        over ``n_streams`` CUDA streams.  A bucket of a few dozen images launches ~26 (resnet18) small kernels whose
        grids fill a fraction of the 148 SMs; running buckets side by side fills the machine.  With ``use_graphs``
        the launch sequence of every (batch, width) shape is captured once into a CUDA graph with its own workspace
        and static input / output buffers, so a bucket costs one graph launch on the host instead of ~26 kernel
        launches and tensor-map encodes.  Results are identical to ``forward`` (same kernels, same data); the
        returned tensors are owned by the graphs and are overwritten by the next call with the same shape."""
        if not batches:
            return []
        dev = batches[0].device
        require_cuda(batches[0], "ResNetEncoder.forward_buckets")
        if getattr(self, "_bucket_streams", None) is None or len(self._bucket_streams) != n_streams:
            self._bucket_streams = [torch.cuda.Stream(dev) for _ in range(n_streams)]
            self._bucket_ws = [Workspace() for _ in range(n_streams)]
        if getattr(self, "_bucket_graphs", None) is None:
            self._bucket_graphs = {}
        with torch.cuda.device(dev):
            self._ensure_packed(dev)
        key_w = self._packed_key
        cur = torch.cuda.current_stream(dev)
        if use_graphs:
            seen = set()
            for x in batches:                    # capture the shapes that are new (or whose weights changed), serially
                k = (tuple(x.shape), x.dtype)
                if k in seen:
                    raise RuntimeError("forward_buckets(use_graphs=True) needs one batch per (shape, dtype)")
                seen.add(k)
                g = self._bucket_graphs.get(k)
                if g is None or g["weights"] != key_w:
                    self._bucket_graphs[k] = self._capture_bucket(x, key_w)
        start = torch.cuda.Event()
        start.record(cur)
        outs = []
        for i, x in enumerate(batches):
            st = self._bucket_streams[i % n_streams]
            if i < n_streams:
                st.wait_event(start)
            with torch.cuda.stream(st):
                if use_graphs:
                    g = self._bucket_graphs[(tuple(x.shape), x.dtype)]
                    g["x"].copy_(x, non_blocking=True)
                    g["graph"].replay()
                    out = g["out"]
                else:
                    out = self.forward(x, _ws=self._bucket_ws[i % n_streams])
                    out.record_stream(cur)
            outs.append(out)
        for st in self._bucket_streams[: min(n_streams, len(batches))]:
            done = torch.cuda.Event()
            done.record(st)
            cur.wait_event(done)
        return outs

    def _capture_bucket(self, x: torch.Tensor, key_w) -> dict:
        """One CUDA graph per (shape, dtype): static input, private workspace, static output."""
        dev = x.device
        ws = Workspace()
        xs = torch.empty_like(x)
        side = torch.cuda.Stream(dev)
        side.wait_stream(torch.cuda.current_stream(dev))
        with torch.cuda.stream(side):
            xs.copy_(x)
            self.forward(xs, _ws=ws)             # warm-up: allocates the workspace, sets function attributes
        torch.cuda.current_stream(dev).wait_stream(side)
        torch.cuda.synchronize(dev)
        graph = torch.cuda.CUDAGraph()
        with torch.cuda.graph(graph):
            out = self.forward(xs, _ws=ws)
        return {"graph": graph, "x": xs, "out": out, "ws": ws, "weights": key_w}

    def forward(self, x: torch.Tensor, _ws: Optional[Workspace] = None) -> torch.Tensor:
        """(B, 3, H, W) -> (B, embedding_dim); any W (the trunk ends in adaptive average
        pooling), which is what width bucketing relies on.  reference encoder.py:231-249."""
        require_cuda(x, "ResNetEncoder.forward")
        if self.training and torch.is_grad_enabled() and any(p.requires_grad for p in self.parameters()):
            raise RuntimeError("ResNetEncoder.forward: only eval-mode inference runs on the native path (BatchNorm is "
                               "folded from the running statistics and no autograd graph is built); call .eval() or "
                               "wrap the call in torch.no_grad()")
        if x.dim() != 4 or x.shape[1] != 3 or x.shape[2] != self.img_height:
            raise RuntimeError(f"ResNetEncoder expected (B,3,{self.img_height},W), got {tuple(x.shape)}")
        with torch.cuda.device(x.device):
            self._ensure_packed(x.device)
            lib = N.lib()
            d = self._desc()
            x = f32c(x)
            B, W = x.shape[0], x.shape[3]
            wsb = lib.i2l_resnet_workspace_bytes(C.byref(d), max(B, 1), W)
            ws = (_ws or self._ws).get(wsb, x.device)
            return torch.ops.i2l.resnet_encoder_fwd(x, self._packed, ws, [d.depth, d.img_height, d.embedding_dim,
                                                                           d.precision])
