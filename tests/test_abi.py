"""CPU suite: the C-ABI shared library loads, exports every symbol include/i2l_b200.h
declares, the ctypes table covers exactly those, and compute entry points FAIL LOUDLY
without a GPU (no fallback)."""
import ctypes as C
import os
import re

import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def declared_symbols():
    src = open(os.path.join(ROOT, "include", "i2l_b200.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(i2l_[a-z0-9_]+)\s*\(", src)))


def test_header_symbols_exported(pkg):
    N = pkg._native
    lib = N.lib()
    decl = declared_symbols()
    assert len(decl) >= 20
    for s in decl:
        assert hasattr(lib, s), f"{s} declared in include/i2l_b200.h but not exported"
    assert sorted(N.SIGNATURES) == decl, "ctypes table and header disagree"
    assert b"sm_100a" in lib.i2l_version()


def test_struct_layouts_match_header(pkg):
    N = pkg._native
    assert C.sizeof(N.CnnDesc) == 4 * (3 + 1 + 8 + 4)
    assert C.sizeof(N.DecDesc) == 24 and C.sizeof(N.ResnetDesc) == 16
    assert C.sizeof(N.CnnParams) == 8 * (8 + 8 + 2)
    assert C.sizeof(N.DecParams) == 8 * (1 + 4 * 8 + 2)
    assert C.sizeof(N.ResnetParams) == 8 + 8 * (5 * 160 + 2)


def test_size_queries_work_without_gpu(pkg):
    N = pkg._native
    lib = N.lib()
    d = N.DecDesc(512, 256, 256, 1, 1, N.FP32)
    assert lib.i2l_dec_packed_bytes(C.byref(d)) > 4 * (512 * 256 + 1024 * 768)
    assert lib.i2l_dec_workspace_bytes(C.byref(d), 32, 150) > 0
    assert lib.i2l_resnet_num_convs(18) == 20 and lib.i2l_resnet_num_convs(50) == 53
    assert lib.i2l_resnet_num_convs(19) < 0


@pytest.mark.skipif(torch.cuda.is_available(), reason="checks the no-GPU failure mode")
def test_compute_fails_loudly_without_gpu(pkg):
    N = pkg._native
    lib = N.lib()
    assert lib.i2l_device_check() == -3
    d = N.DecDesc(46, 32, 32, 1, 1, N.FP32)
    rc = lib.i2l_decode_greedy(C.byref(d), None, None, 1, 1, 2, 5, 1.0, 1, None, None, None, None, 0, None)
    assert rc == -3 and b"no CPU fallback" in lib.i2l_last_error()
    m = pkg.Seq2SeqModel("cnn_lstm", 46, dict(img_height=16, img_width=32, channels=1, embedding_dim=32,
                                              conv_filters=[4]), dict(hidden_dim=32, attention=True))
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        m.encoder(torch.zeros(1, 1, 16, 32))
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        m.decoder.decode_step(torch.zeros(1, 32), torch.zeros(1, 1, dtype=torch.long))


def test_state_dict_keys_match_reference(pkg):
    """the drop-in modules accept the reference's checkpoints."""
    m = pkg.Seq2SeqModel("cnn_lstm", 46, dict(img_height=16, img_width=32, channels=3, embedding_dim=32),
                         dict(hidden_dim=32, lstm_layers=2, attention=True))
    keys = set(m.state_dict())
    for k in ["encoder.cnn_layers.0.weight", "encoder.cnn_layers.3.bias", "encoder.cnn_layers.6.weight",
              "encoder.embedding_layer.weight", "decoder.embedding.weight", "decoder.lstm.weight_ih_l0",
              "decoder.lstm.weight_hh_l1", "decoder.lstm.bias_ih_l1", "decoder.attention.attn.weight",
              "decoder.attention.attn.bias", "decoder.attention.v.weight", "decoder.output_layer.weight",
              "decoder.output_layer.bias"]:
        assert k in keys
    r = pkg.Seq2SeqModel("resnet_lstm", 46, dict(img_height=64, img_width=96, model_name="resnet18", embedding_dim=32),
                         dict(hidden_dim=32, attention=True))
    rk = set(r.state_dict())
    for k in ["encoder.resnet.0.weight", "encoder.resnet.1.running_mean", "encoder.resnet.4.0.conv1.weight",
              "encoder.resnet.5.0.downsample.0.weight", "encoder.resnet.7.1.bn2.num_batches_tracked"]:
        assert k in rk
    with pytest.raises(ValueError, match="Invalid model type"):
        pkg.Seq2SeqModel("vit_lstm", 46)
    with pytest.raises(ValueError, match="Invalid ResNet model name"):
        pkg.ResNetEncoder(model_name="resnet19")


def test_cli_predict_surface(pkg):
    """`img2latex predict` option surface (reference cli.py:253-270): same positionals, option names and defaults."""
    from hmer_img2latex_b200.cli import build_parser, main
    a = build_parser().parse_args(["predict", "ckpt.pt", "img.png"])
    assert (a.beam_size, a.max_length, a.temperature, a.top_k, a.top_p, a.device) == (0, 141, 1.0, 0, 0.0, None)
    a = build_parser().parse_args(["predict", "c", "i", "--beam-size", "5", "--max-length", "50", "--temperature", "0.8",
                                   "--top-k", "40", "--top-p", "0.9", "--device", "cuda:0"])
    assert (a.beam_size, a.max_length, a.temperature, a.top_k, a.top_p, a.device) == (5, 50, 0.8, 40, 0.9, "cuda:0")
    with pytest.raises(RuntimeError, match="CUDA"):
        main(["predict", "c", "i", "--device", "cpu"])
