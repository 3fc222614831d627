"""GPU suite: the CUDA path (fp32 mode, through the C-ABI) against the golden vectors the
LIVE reference produced (tests/golden/).  Tolerance 1e-3 relative for floats (north_star);
token sequences identical (the golden cases contain no near ties)."""
import numpy as np
import pytest
import torch

import helpers as H
from helpers import oracle
from test_oracle_golden import PB_CFG, load, pb_inputs, unpad

pytestmark = pytest.mark.gpu
TOL = 1e-3


def cu(a):
    return torch.as_tensor(np.asarray(a)).cuda()


@pytest.mark.parametrize("name,cfg", [("cnn_headline_sharp.npz", H.HEADLINE), ("cnn_headline_default.npz", H.HEADLINE),
                                      ("cnn_small_l2_beam.npz", H.SMALL)])
def test_seq2seq_vs_reference_golden(pkg, name, cfg):
    d = load(name)
    B, T = int(d["B"]), int(d["T"])
    p = oracle.make_params(cfg, int(d["seed"]), sharp=bool(d["sharp"]))
    m = H.build_model(pkg, cfg, p)
    x = H.make_images(cfg, B).cuda()
    enc = m.encoder(x)
    assert H.rel_err(enc, torch.as_tensor(d["enc"])) < TOL
    l1, (h1, c1) = m.decoder.decode_step(enc, cu(d["tok"]), None)
    assert H.rel_err(l1, torch.as_tensor(d["logits1"])) < TOL and H.rel_err(c1, torch.as_tensor(d["c1"])) < TOL
    l2, (h2, c2) = m.decoder.decode_step(enc, cu(d["tok2"]), (cu(d["h1"]), cu(d["c1"])))
    assert H.rel_err(l2, torch.as_tensor(d["logits2"])) < TOL and H.rel_err(h2, torch.as_tensor(d["h2"])) < TOL
    assert m.inference(x, H.START, H.END, max_length=T) == unpad(d["greedy_raw"])
    assert [m.inference(x[i:i + 1], H.START, H.END, max_length=T) for i in range(B)] == unpad(d["greedy_single"])
    if "beam" in d:
        K = int(d["beam_size"])
        assert [m.inference(x[i:i + 1], H.START, H.END, max_length=T, beam_size=K) for i in range(B)] == unpad(d["beam"])
        assert m.beam_search_batch(enc, H.START, H.END, T, K) == unpad(d["beam"])


@pytest.mark.parametrize("name,cfg", [("resnet18.npz", H.R18), ("resnet50.npz", H.R50)])
def test_resnet_vs_reference_golden(pkg, name, cfg):
    d = load(name)
    p = oracle.make_params(cfg, 0)
    m = H.build_model(pkg, cfg, p)
    for w in d["widths"]:
        x = H.make_images(cfg, 2, width=int(w)).cuda()
        assert H.rel_err(m.encoder(x), torch.as_tensor(d[f"enc_w{w}"])) < TOL


def test_attention_vs_reference_golden(pkg):
    d = load("attention_L5.npz")
    att = pkg.Attention(48, 32)
    att.load_state_dict({"attn.weight": torch.as_tensor(d["w"]), "attn.bias": torch.as_tensor(d["b"]),
                         "v.weight": torch.as_tensor(d["v"])})
    out = att.cuda()(cu(d["hid"]), cu(d["enc"]))
    assert H.rel_err(out, torch.as_tensor(d["ctx"])) < TOL


@pytest.mark.parametrize("name", ["predict_batch_greedy.npz", "predict_batch_topk_topp.npz", "predict_batch_topp.npz"])
def test_predict_batch_vs_reference_golden(pkg, name):
    d = load(name)
    p, x = pb_inputs(d)
    m = H.build_model(pkg, PB_CFG, p)
    tok = pkg.LaTeXTokenizer(); tok.default_init()
    pred = pkg.Predictor(m, tok)
    T = int(d["T"])
    got = pred.predict_batch(list(x), max_length=T, temperature=float(d["temperature"]), top_k=int(d["top_k"]),
                             top_p=float(d["top_p"]), batch_size=int(d["B"]), uniforms=torch.as_tensor(d["u"]))
    assert got == [str(s) for s in d["strings"]]


def test_from_checkpoint_vs_reference_golden(pkg, tmp_path):
    """a checkpoint file in the reference's layout (training/trainer.py:207-221) loaded by
    Predictor.from_checkpoint gives the strings the reference's Predictor.from_checkpoint(...).predict /
    predict_batch produced from the same file contents (tests/golden/checkpoint.npz)."""
    d = load("checkpoint.npz")
    p = oracle.make_params(H.CKPT_CFG, 2, sharp=True)
    path = str(tmp_path / "best_checkpoint_epoch_3_step_7.pt")
    H.make_checkpoint(path, p)
    pred = pkg.Predictor.from_checkpoint(path)
    assert pred.tokenizer.vocab_size == int(d["vocab_size"]) and pred.tokenizer.max_sequence_length == 30
    assert pred.model.decoder.lstm_layers == 2 and pred.model.encoder.img_width == 800
    g = torch.Generator().manual_seed(13)
    imgs = [torch.rand(1, 64, 800, generator=g) for _ in range(4)]
    assert pred.predict_batch(imgs, max_length=24, batch_size=4) == d["batch"].tolist()
    assert [pred.predict(im, max_length=24) for im in imgs] == d["singles"].tolist()


def test_cli_predict_on_reference_checkpoint(pkg, tmp_path, capsys):
    """`img2latex predict CKPT IMG` end to end: checkpoint file in the trainer's layout + PNG on disk -> the string
    Predictor.predict gives for the same path (device-side load_image: convert L, LANCZOS resize, pad, x/255*2-1)."""
    from PIL import Image
    from hmer_img2latex_b200.cli import main
    p = oracle.make_params(H.CKPT_CFG, 2, sharp=True)
    ck = str(tmp_path / "best_checkpoint_epoch_3_step_7.pt")
    H.make_checkpoint(ck, p)
    rng = np.random.default_rng(5)
    img = np.full((40, 300, 3), 255, np.uint8)
    for _ in range(40):
        y, x = int(rng.integers(0, 36)), int(rng.integers(0, 290))
        img[y:y + 3, x:x + 8] = 0
    path = str(tmp_path / "formula.png")
    Image.fromarray(img, "RGB").save(path)
    assert main(["predict", ck, path, "--max-length", "24"]) == 0
    out = capsys.readouterr().out.strip().splitlines()
    assert out[0] == "Generated LaTeX:"
    pred = pkg.Predictor.from_checkpoint(ck)
    assert out[1] == pred.predict(path, max_length=24)
    # the oracle on the reference's load_image arithmetic gives the same ids
    g = oracle.resize.resize_with_aspect_ratio(oracle.resize.rgb_to_l(img), 64, 800)
    x = oracle.normalize_u8(torch.from_numpy(g)[None, None], "pm1")
    enc = oracle.encoder(p, x, H.CKPT_CFG)
    ids = oracle.inference_postprocess(oracle.greedy_search(p, enc, H.START, H.END, 24, 1.0, H.CKPT_CFG)[0], H.START, H.END)
    ids = ids[0] if ids and isinstance(ids[0], list) else ids
    assert out[1] == pred.tokenizer.decode([i for i in ids if i != H.END])
