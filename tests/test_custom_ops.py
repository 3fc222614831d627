"""CPU suite: the `torch.library` operator layer (hmer-img2latex_b200/ops.py) -- north_star / SURVEY 8b: the
reference's modules are swapped for PyTorch custom ops that call the sm_100a kernels through the C-ABI.  Without a
GPU we check the registration (names, CUDA-only dispatch: no CPU kernel exists), the fake (meta) kernels' shapes and
dtypes, and that the drop-in modules really route through the operators."""
import pytest
import torch


def test_ops_are_registered_cuda_only(pkg):
    for name in pkg.ops.OPS:
        assert hasattr(torch.ops.i2l, name), name
        op = getattr(torch.ops.i2l, name).default
        assert torch._C._dispatch_has_kernel_for_dispatch_key(op.name(), "CUDA"), name
        assert not torch._C._dispatch_has_kernel_for_dispatch_key(op.name(), "CPU"), f"{name}: a CPU kernel exists"


def test_cpu_tensors_fail_in_the_dispatcher(pkg):
    with pytest.raises((NotImplementedError, RuntimeError)):
        torch.ops.i2l.decode_greedy(torch.zeros(2, 32), torch.zeros(8, dtype=torch.uint8), torch.zeros(8, dtype=torch.uint8),
                                    [46, 32, 32, 1, 1, 0], 1, 2, 5, 1.0, 1)


def test_fake_kernels_propagate_shapes(pkg):
    """Shapes / dtypes under FakeTensorMode are those of the real kernels (what torch.compile / export trace)."""
    from torch._subclasses.fake_tensor import FakeTensorMode
    with FakeTensorMode():
        dev = "cuda"
        u8 = lambda n: torch.empty(n, dtype=torch.uint8, device=dev)
        B, V, E, Hd, L, T, K = 6, 512, 256, 256, 1, 20, 5
        dd = [V, E, Hd, L, 1, 1]
        x = torch.empty(B, 3, 64, 320, device=dev)
        cd = [64, 320, 3, 3, 2, E, 1, 32, 64, 128]
        enc = torch.ops.i2l.cnn_encoder_fwd(x, u8(16), u8(16), cd)
        assert enc.shape == (B, E) and enc.dtype == torch.float32 and enc.device.type == "cuda"
        enc8 = torch.ops.i2l.cnn_encoder_fwd_u8(x.to(torch.uint8), u8(16), u8(16), cd, 0, [0.5] * 3, [0.5] * 3)
        assert enc8.shape == (B, E)
        assert torch.ops.i2l.resnet_encoder_fwd(x, u8(16), u8(16), [50, 64, E, 1]).shape == (B, E)
        ctx = torch.ops.i2l.attention_fwd(torch.empty(B, Hd, device=dev), torch.empty(B, 7, E, device=dev),
                                          torch.empty(Hd, Hd + E, device=dev), torch.empty(Hd, device=dev),
                                          torch.empty(1, Hd, device=dev), u8(16))
        assert ctx.shape == (B, E)
        tok = torch.empty(B, dtype=torch.int64, device=dev)
        lg, h, c, bad = torch.ops.i2l.decode_step(enc, tok, None, None, u8(16), u8(16), dd)
        assert lg.shape == (B, V) and h.shape == (L, B, Hd) and c.shape == (L, B, Hd) and bad.dtype == torch.int32
        lg, h, c, bad = torch.ops.i2l.decoder_forward(enc, torch.empty(B, T, dtype=torch.int64, device=dev), None, None,
                                                      u8(16), u8(16), dd)
        assert lg.shape == (B, T, V)
        t, ln, st = torch.ops.i2l.decode_greedy(enc, u8(16), u8(16), dd, 1, 2, T, 1.0, 1)
        assert t.shape == (B, T + 1) and t.dtype == torch.int64 and ln.shape == (B,) and st.shape == ()
        t, ln, st, pr = torch.ops.i2l.decode_sample(enc, u8(16), u8(16), dd, 1, 2, T, 0.8, 50, 0.9, 0, 0, None, True)
        assert pr.shape == (T, B, V)
        o = torch.ops.i2l.decode_beam(enc, u8(16), u8(16), dd, K, 1, 2, T, True, True)
        assert o[0].shape == (B, T) and o[2].dtype == torch.float64 and o[3].shape == (T, B, K) and o[6].shape == (T, B, K, K)
        o = torch.ops.i2l.decode_beam(enc, u8(16), u8(16), dd, K, 1, 2, T, False, False)
        assert o[3].numel() == 0 and o[7].numel() == 0


def test_modules_route_through_the_ops(pkg):
    """The drop-in modules call torch.ops.i2l.* (and nothing else) for their compute: source-level check of the
    module files -- no direct lib.i2l_*_fwd / i2l_decode_* ctypes calls are left in model/."""
    import inspect
    import re
    for mod in (pkg.model.encoder, pkg.model.decoder):
        src = inspect.getsource(mod)
        assert "torch.ops.i2l." in src
        direct = re.findall(r"lib\.(i2l_(?:cnn_encoder_fwd\w*|resnet_encoder_fwd|attention_fwd|decode_\w+|decoder_forward))\(", src)
        assert not direct, direct
