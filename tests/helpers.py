"""Shared test helpers: configurations, seeded inputs, parameter loading."""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

import oracle  # noqa: E402  (test infrastructure)

START, END = 1, 2

HEADLINE = dict(model_type="cnn_lstm", vocab_size=512, embedding_dim=256, hidden_dim=256, lstm_layers=1,
                attention=True, img_height=64, img_width=320, channels=3)
SMALL = dict(model_type="cnn_lstm", vocab_size=46, embedding_dim=32, hidden_dim=48, lstm_layers=2,
             attention=True, img_height=16, img_width=40, channels=1, conv_filters=[8, 16])
# the reference's shipped decoder (img2latex/configs/config.yaml:45-50): E = H = 512, two LSTM layers
SHIPPED = dict(model_type="cnn_lstm", vocab_size=512, embedding_dim=512, hidden_dim=512, lstm_layers=2, attention=True,
               img_height=16, img_width=32, channels=1, conv_filters=[8])
R18 = dict(model_type="resnet_lstm", model_name="resnet18", vocab_size=46, embedding_dim=64, hidden_dim=64,
           lstm_layers=1, attention=True, img_height=64, img_width=128, channels=3)
R50 = dict(model_type="resnet_lstm", model_name="resnet50", vocab_size=46, embedding_dim=64, hidden_dim=64,
           lstm_layers=1, attention=True, img_height=64, img_width=96, channels=3)


def make_images(cfg, batch, seed=1, width=None):
    """Seeded synthetic images: randn plus a per-image gain / offset so encodings differ."""
    g = torch.Generator().manual_seed(seed)
    w = cfg["img_width"] if width is None else width
    x = torch.randn(batch, cfg.get("channels", 3), cfg["img_height"], w, generator=g)
    gain = 0.5 + torch.rand(batch, 1, 1, 1, generator=g)
    off = torch.randn(batch, 1, 1, 1, generator=g) * 0.5
    return x * gain + off


def build_model(pkg, cfg, params, precision="fp32", device="cuda"):
    enc = {k: cfg[k] for k in ("img_height", "img_width", "channels", "embedding_dim", "conv_filters",
                               "kernel_size", "pool_size", "model_name") if k in cfg}
    dec = dict(hidden_dim=cfg.get("hidden_dim", 256), lstm_layers=cfg.get("lstm_layers", 1),
               attention=cfg.get("attention", True))
    m = pkg.Seq2SeqModel(cfg.get("model_type", "cnn_lstm"), cfg["vocab_size"], enc, dec, precision=precision)
    missing, unexpected = m.load_state_dict(params, strict=True)
    assert not missing and not unexpected
    return m.to(device).eval()


def rel_err(a, b):
    a, b = a.double().cpu(), b.double().cpu()
    return float((a - b).abs().max() / b.abs().max().clamp_min(1e-30))


def default_vocab():
    """The 42 non-special tokens of the reference tokenizer's `default_init` (data/tokenizer.py:323-385), in id order."""
    import i2l_import
    t = i2l_import.load().LaTeXTokenizer()
    t.default_init()
    return [t.id_to_token[i] for i in range(4, t.vocab_size)]


CKPT_CFG = dict(model_type="cnn_lstm", vocab_size=46, embedding_dim=32, hidden_dim=48, lstm_layers=2, attention=True,
                img_height=64, img_width=800, channels=1, conv_filters=[4, 8, 8])


def make_checkpoint(path, params):
    """A checkpoint in the layout training/trainer.py:207-221 writes (config in configs/config.yaml's layout)."""
    ref_tok_ids = {"<PAD>": 0, "<START>": 1, "<END>": 2, "<UNK>": 3}
    for t in default_vocab():
        ref_tok_ids[t] = len(ref_tok_ids)
    c = CKPT_CFG
    config = {"model": {"name": "cnn_lstm", "embedding_dim": c["embedding_dim"],
                        "encoder": {"cnn": dict(img_height=64, img_width=800, channels=1, conv_filters=c["conv_filters"],
                                                kernel_size=3, pool_size=2, padding="same"),
                                    "resnet": dict(img_height=64, img_width=800, channels=3, model_name="resnet18")},
                        "decoder": dict(hidden_dim=c["hidden_dim"], lstm_layers=c["lstm_layers"], dropout=0.3, attention=True)}}
    torch.save({"epoch": 3, "step": 7, "model_state_dict": params, "optimizer_state_dict": {}, "metrics": {},
                "config": config,
                "tokenizer_config": {"token_to_id": ref_tok_ids,
                                     "special_tokens": {"PAD": "<PAD>", "START": "<START>", "END": "<END>", "UNK": "<UNK>"},
                                     "max_sequence_length": 30}}, path)
