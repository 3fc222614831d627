"""Import shim for the LIVE reference (``/root/reference``), build-container only.

TEST INFRASTRUCTURE -- see ``oracle/__init__.py``.  Used by
``tests/golden/make_golden.py`` to pin the oracle, and by the optional
``tests/test_oracle_vs_live_reference.py`` (skipped where the reference tree is
absent, e.g. on the GPU box).  Recipe from SURVEY.md F10/F11:
  * stub ``matplotlib`` (imported by img2latex/utils/visualize_metrics.py:10),
  * rebind torchvision resnet constructors to ``weights=None``
    (img2latex/model/encoder.py:185-194 always asks for ImageNet weights).
"""
from __future__ import annotations

import os
import sys
import types

_TRAVELLING = os.path.join(os.path.dirname(os.path.abspath(__file__)), "_ref")   # oracle/make_ref.py


def _default_root() -> str:
    """/root/reference in the build container, else the copy oracle/make_ref.py left under oracle/_ref/."""
    if os.path.isdir("/root/reference/img2latex/model"):
        return "/root/reference"
    return _TRAVELLING


REFERENCE_ROOT = os.environ.get("I2L_REFERENCE_ROOT") or _default_root()


def available() -> bool:
    return os.path.isdir(os.path.join(REFERENCE_ROOT, "img2latex", "model"))


def load():
    """Returns the reference's (Seq2SeqModel, LSTMDecoder, Attention, CNNEncoder,
    ResNetEncoder, Predictor, LaTeXTokenizer) classes."""
    if not available():
        raise RuntimeError(f"reference tree not found at {REFERENCE_ROOT}")
    sys.dont_write_bytecode = True
    for name in ("matplotlib", "matplotlib.pyplot", "seaborn"):
        if name not in sys.modules:
            try:
                __import__(name)
            except Exception:
                sys.modules[name] = types.ModuleType(name)
    if "matplotlib.pyplot" in sys.modules and not hasattr(sys.modules["matplotlib"], "pyplot"):
        sys.modules["matplotlib"].pyplot = sys.modules["matplotlib.pyplot"]
    if REFERENCE_ROOT not in sys.path:
        sys.path.insert(0, REFERENCE_ROOT)
    import logging
    logging.disable(logging.INFO)
    import img2latex.model.encoder as enc_mod
    import torchvision.models as tvm

    class _Models:
        def __getattr__(self, k):
            return getattr(tvm, k)

    shim = _Models()
    for nm in ("resnet18", "resnet34", "resnet50", "resnet101", "resnet152"):
        ctor = getattr(tvm, nm)
        object.__setattr__(shim, nm, (lambda c: (lambda weights=None, **kw: c(weights=None, **kw)))(ctor))
    enc_mod.models = shim
    from img2latex.model.seq2seq import Seq2SeqModel
    from img2latex.model.decoder import LSTMDecoder, Attention
    from img2latex.model.encoder import CNNEncoder, ResNetEncoder
    from img2latex.training.predictor import Predictor
    from img2latex.data.tokenizer import LaTeXTokenizer
    return dict(Seq2SeqModel=Seq2SeqModel, LSTMDecoder=LSTMDecoder, Attention=Attention,
                CNNEncoder=CNNEncoder, ResNetEncoder=ResNetEncoder, Predictor=Predictor,
                LaTeXTokenizer=LaTeXTokenizer)


def reference_module(name: str):
    """Import a module of the live reference by dotted name (after the shims of `load`)."""
    load()
    import importlib
    return importlib.import_module(name)


def build_reference_model(cfg: dict, params: dict):
    """Instantiate the reference Seq2SeqModel for ``cfg`` and load ``params``
    (keys = reference state_dict keys) into it; eval mode."""
    ref = load()
    enc = dict(img_height=cfg["img_height"], img_width=cfg["img_width"], channels=cfg.get("channels", 3),
               embedding_dim=cfg.get("embedding_dim", 256))
    if cfg.get("model_type", "cnn_lstm") == "cnn_lstm":
        enc.update(conv_filters=cfg.get("conv_filters", [32, 64, 128]), kernel_size=cfg.get("kernel_size", 3),
                   pool_size=cfg.get("pool_size", 2))
    else:
        enc.update(model_name=cfg.get("model_name", "resnet50"))
    dec = dict(hidden_dim=cfg.get("hidden_dim", 256), lstm_layers=cfg.get("lstm_layers", 1),
               attention=cfg.get("attention", True), max_seq_length=cfg.get("max_seq_length", 150))
    m = ref["Seq2SeqModel"](cfg.get("model_type", "cnn_lstm"), cfg["vocab_size"], enc, dec)
    missing, unexpected = m.load_state_dict(params, strict=True)
    assert not missing and not unexpected
    return m.eval()
