"""CPU suite: pins the oracle (oracle/port.py) against the golden vectors produced by the
LIVE reference (tests/golden/make_golden.py).  Tokens / strings must be identical; float
tensors agree to 1e-6 (same ATen CPU kernels, possibly different summation grouping where
the port writes out nn.LSTM by hand)."""
import os

import numpy as np
import pytest
import torch

import helpers as H
from helpers import oracle

G = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
torch.set_num_threads(min(8, os.cpu_count() or 1))


def load(name):
    return np.load(os.path.join(G, name), allow_pickle=False)


def unpad(rows):
    return [[int(t) for t in r if t >= 0] for r in rows]


def close(a, b, tol=2e-6):
    a, b = torch.as_tensor(np.asarray(a)).double(), torch.as_tensor(np.asarray(b)).double()
    assert a.shape == b.shape
    assert float((a - b).abs().max()) <= tol * max(1.0, float(b.abs().max()))


def check_inputs(d, p, x, key="checksum"):
    s = sum(float(v.double().abs().sum()) for v in p.values() if v.dtype.is_floating_point)
    assert abs(s - d[key][0]) < 1e-6 * abs(d[key][0]) and abs(float(x.double().abs().sum()) - d[key][1]) < 1e-6 * d[key][1], \
        "seeded inputs drifted from the ones the golden vectors were generated with"


@pytest.mark.parametrize("name,cfg", [("cnn_headline_sharp.npz", H.HEADLINE), ("cnn_headline_default.npz", H.HEADLINE),
                                      ("cnn_small_l2_beam.npz", H.SMALL)])
def test_seq2seq_golden(name, cfg):
    d = load(name)
    B, T = int(d["B"]), int(d["T"])
    p = oracle.make_params(cfg, int(d["seed"]), sharp=bool(d["sharp"]))
    x = H.make_images(cfg, B)
    check_inputs(d, p, x)
    with torch.no_grad():
        enc = oracle.encoder(p, x, cfg)
        close(enc, d["enc"])
        l1, (h1, c1) = oracle.decode_step(p, enc, torch.as_tensor(d["tok"]), None, cfg)
        close(l1, d["logits1"]); close(h1, d["h1"]); close(c1, d["c1"])
        l2, (h2, c2) = oracle.decode_step(p, enc, torch.as_tensor(d["tok2"]), (h1, c1), cfg)
        close(l2, d["logits2"]); close(h2, d["h2"]); close(c2, d["c2"])
        raw, _ = oracle.greedy_search(p, enc, H.START, H.END, T, 1.0, cfg)
        assert raw == unpad(d["greedy_raw"])
        single = [oracle.inference_postprocess(oracle.greedy_search(p, enc[i:i + 1], H.START, H.END, T, 1.0, cfg)[0],
                                               H.START, H.END) for i in range(B)]
        assert single == unpad(d["greedy_single"])
        if "beam" in d:
            got = oracle.beam_search_batched(p, enc, H.START, H.END, T, int(d["beam_size"]), cfg)
            assert got == unpad(d["beam"])


@pytest.mark.parametrize("name,cfg", [("resnet18.npz", H.R18), ("resnet50.npz", H.R50)])
def test_resnet_golden(name, cfg):
    d = load(name)
    p = oracle.make_params(cfg, 0)
    for w in d["widths"]:
        x = H.make_images(cfg, 2, width=int(w))
        check_inputs(d, p, x, f"checksum_w{w}")
        with torch.no_grad():
            close(oracle.resnet_encoder(p, x, cfg["model_name"]), d[f"enc_w{w}"], tol=1e-5)


def test_attention_golden():
    d = load("attention_L5.npz")
    t = {k: torch.as_tensor(d[k]) for k in d.files}
    close(oracle.attention(t["w"], t["b"], t["v"], t["hid"], t["enc"]), t["ctx"])
    # src_len == 1 is the identity, bit for bit
    one = oracle.attention(t["w"], t["b"], t["v"], t["hid"], t["enc"][:, :1])
    assert torch.equal(one, t["enc"][:, :1])


PB_CFG = dict(model_type="cnn_lstm", vocab_size=46, embedding_dim=32, hidden_dim=48, lstm_layers=1, attention=True,
              img_height=64, img_width=800, channels=1, conv_filters=[4, 8, 8])


def pb_inputs(d):
    B = int(d["B"])
    p = oracle.make_params(PB_CFG, 3, sharp=True)
    g = torch.Generator().manual_seed(11)
    x = torch.stack([torch.rand(1, 64, 800, generator=g) for _ in range(B)])
    return p, x


@pytest.mark.parametrize("name", ["predict_batch_greedy.npz", "predict_batch_topk_topp.npz", "predict_batch_topp.npz"])
def test_predict_batch_golden(name, pkg):
    d = load(name)
    p, x = pb_inputs(d)
    T = int(d["T"])
    with torch.no_grad():
        enc = oracle.encoder(p, x, PB_CFG)
        seqs, trimmed, steps, ptrace = oracle.sample_loop(p, enc, H.START, H.END, T, float(d["temperature"]),
                                                          int(d["top_k"]), float(d["top_p"]), PB_CFG,
                                                          uniforms=torch.as_tensor(d["u"]), return_probs=True)
    tok = pkg.LaTeXTokenizer(); tok.default_init()
    strs = []
    for s in trimmed:                                   # predictor.py:382-392
        s = s[1:] if s and s[0] == H.START else s
        s = s[:-1] if s and s[-1] == H.END else s
        strs.append(tok.decode(s))
    assert strs == [str(s) for s in d["strings"]]
    if "probs" in d:
        assert len(ptrace) == d["probs"].shape[0]
        close(torch.stack(ptrace), d["probs"])


def test_tokenizer_default_vocab(pkg):
    tok = pkg.LaTeXTokenizer(); tok.default_init()
    assert tok.vocab_size == 46 and (tok.pad_token_id, tok.start_token_id, tok.end_token_id, tok.unk_token_id) == (0, 1, 2, 3)
    assert tok.decode([1, 13, 4, 99, 2]) == "\\frac + <UNK>"


def test_filter_probs_edge_cases():
    lg = torch.tensor([[0.0, 0.0, 0.0, 0.0], [5.0, 1.0, 1.0, -2.0]])
    pr = oracle.filter_probs(lg, 1.0, 2, 0.0)           # ties at the k-th value are all kept (predictor.py:305)
    assert torch.allclose(pr[0], torch.full((4,), 0.25)) and int((pr[1] > 0).sum()) == 3
    pr = oracle.filter_probs(lg, 1.0, 0, 0.5)            # nucleus always keeps sorted index 0 (predictor.py:320)
    assert int((pr[1] > 0).sum()) == 1 and abs(float(pr[1].sum()) - 1) < 1e-6
    pr = oracle.filter_probs(lg, 1.0, 100, 0.0)          # top_k clamped to V (predictor.py:300)
    assert torch.allclose(pr, torch.softmax(lg, -1))
    u = torch.tensor([0.0, 0.999999])
    assert oracle.inverse_cdf_draw(torch.tensor([[0.0, 0.5, 0.5, 0.0], [0.2, 0.8, 0.0, 0.0]]), u).tolist() == [1, 1]


def test_load_image_golden():
    """oracle.normalize_u8 == the pixel arithmetic of the live reference's load_image."""
    d = load("load_image.npz")
    rgb = torch.as_tensor(d["rgb"]).permute(2, 0, 1)          # HWC -> CHW (data/utils.py:63-65)
    gray = torch.as_tensor(d["gray"]).unsqueeze(0)
    assert torch.equal(oracle.normalize_u8(rgb, "meanstd"), torch.as_tensor(d["out_rgb"]))
    assert torch.equal(oracle.normalize_u8(gray, "pm1"), torch.as_tensor(d["out_gray"]))


def test_fused_affine_equals_reference_after_bf16():
    """The uint8-input conv1 (cnn_bf16.cu) builds its operand as bf16(fma(x, a, b)); the reference arithmetic is
    x/255*2-1 or (x/255-mean)/std in fp32 (data/utils.py:68-80, predictor.py:441-446).  For all 256 pixel values the
    two agree after rounding to bf16, so the fused path equals normalize_u8 -> bf16 -> conv1 bit for bit."""
    import numpy as np
    x = torch.arange(256, dtype=torch.uint8).reshape(1, 1, 16, 16)
    ref = oracle.normalize_u8(x, "pm1").flatten()
    a = np.float32(2.0) / np.float32(255.0)
    fused = (x.flatten().double() * float(a) - 1.0).float()          # fmaf: exact product, one rounding
    assert torch.equal(ref.bfloat16(), fused.bfloat16())
    rgb = torch.arange(256, dtype=torch.uint8).reshape(1, 1, 16, 16).repeat(1, 3, 1, 1)
    ref3 = oracle.normalize_u8(rgb, "meanstd")
    for c, (m, s) in enumerate(zip((0.485, 0.456, 0.406), (0.229, 0.224, 0.225))):
        m32, s32 = np.float32(m), np.float32(s)
        a = np.float32(1.0) / (np.float32(255.0) * s32)
        b = -m32 / s32
        fused = (x.flatten().double() * float(a) + float(b)).float()
        assert torch.equal(ref3[0, c].flatten().bfloat16(), fused.bfloat16())


@pytest.mark.parametrize("tag,cfg", [("headline", H.HEADLINE), ("small_l2", H.SMALL),
                                     ("small_l2_noattn", dict(H.SMALL, attention=False))])
def test_teacher_forced_golden(tag, cfg):
    """`Seq2SeqModel.forward` of the live reference in eval mode (both decoder branches) against the
    restated recurrence (oracle.seq2seq_forward)."""
    d = load("teacher_forced.npz")
    seed, B, T = (int(v) for v in d[f"{tag}_meta"])
    p = oracle.make_params(cfg, seed, sharp=True)
    x = H.make_images(cfg, B)
    check_inputs(d, p, x, key=f"{tag}_checksum")
    tgt = torch.as_tensor(d[f"{tag}_target"])
    with torch.no_grad():
        out = oracle.seq2seq_forward(p, x, tgt, cfg)
    assert out.shape == (B, T, cfg["vocab_size"])
    close(out, d[f"{tag}_logits"])


def _resize_golden_images(d):
    out, off = [], 0
    for h, w, c in d["shapes"]:
        n = int(h) * int(w) * int(c)
        a = d["pixels"][off:off + n]
        out.append(a.reshape(h, w, c) if c == 3 else a.reshape(h, w))
        off += n
    return out


def test_resize_with_aspect_ratio_golden():
    """oracle/resize.py (Pillow's 8-bit resampler restated) against the reference's ResizeWithAspectRatio run on
    live Pillow: identical bytes."""
    d = load("resize.npz")
    TH, TW = (int(v) for v in d["target"])
    imgs = _resize_golden_images(d)
    got = [oracle.resize.resize_with_aspect_ratio(a, TH, TW) for a in imgs]
    assert np.array_equal(np.stack([g for g in got if g.ndim == 2]), d["out_l"])
    assert np.array_equal(np.stack([g for g in got if g.ndim == 3]), d["out_rgb"])


def test_load_image_geometry_golden():
    """Reference load_image through PNG files (convert -> resize/pad -> /255 -> normalise) restated end to end."""
    d = load("resize.npz")
    TH, TW = (int(v) for v in d["target"])
    for j in range(2):
        rgb = d[f"file{j}"]
        g = oracle.resize.resize_with_aspect_ratio(oracle.resize.rgb_to_l(rgb), TH, TW)
        assert torch.equal(oracle.normalize_u8(torch.from_numpy(g)[None], "pm1"), torch.from_numpy(d[f"file{j}_load1"]))
        c = oracle.resize.resize_with_aspect_ratio(rgb, TH, TW)
        out3 = oracle.normalize_u8(torch.from_numpy(c).permute(2, 0, 1).contiguous(), "meanstd")
        assert torch.equal(out3, torch.from_numpy(d[f"file{j}_load3"]))
    # PIL branch of Predictor._prepare_image: convert L, plain bicubic stretch to 64x800, /255*2-1
    g = oracle.resize.resize_lanczos_u8(oracle.resize.rgb_to_l(d["pil_in"]), 800, 64, "bicubic")
    assert torch.equal(oracle.normalize_u8(torch.from_numpy(g)[None, None], "pm1"), torch.from_numpy(d["pil_prepared"]))


def test_resize_plan_host_tables(pkg):
    """Host half of the C-ABI resize path (i2l_resize_plan_build, no GPU needed): geometry and 22-bit filter
    tables of a ragged batch equal the oracle's restatement of Pillow's precompute_coeffs."""
    P = pkg.preprocess
    d = load("resize.npz")
    TH, TW = (int(v) for v in d["target"])
    imgs = [a for a in _resize_golden_images(d) if a.ndim == 2]
    for resample in ("lanczos", "bicubic"):
        plan = P.ResizePlan(imgs, TH, TW, resample=resample)
        for a, pl in zip(imgs, plan.describe()):
            nw = oracle.resize.aspect_width(a.shape[1], a.shape[0], TH)
            assert (pl["new_w"], pl["new_h"]) == (nw, TH)
            assert pl["left"] == (max(nw - TW, 0) // 2) and pl["kept_w"] == min(nw, TW)
            for tag, n_in, n_out in (("h", a.shape[1], nw), ("v", a.shape[0], TH)):
                if n_in == n_out:
                    assert pl[f"weights_{tag}"] is None          # Resample.c: identity pass is skipped
                    continue
                ks, bnd, kk = oracle.resize.lanczos_coeffs(n_in, n_out, resample)
                assert np.array_equal(pl[f"bounds_{tag}"], bnd)
                assert np.array_equal(pl[f"weights_{tag}"][:, :ks], kk) and not pl[f"weights_{tag}"][:, ks:].any()
    with pytest.raises(ValueError):                               # Pillow: "height and width must be > 0"
        P.ResizePlan([np.zeros((200, 1), np.uint8)], 16, 100)
    white = P.ResizePlan([np.zeros((0, 7), np.uint8)], 16, 100).describe()[0]
    assert white["kept_w"] == 0                                   # transforms.py:28-29: all-white output


def test_metrics_golden():
    """oracle/metrics.py against the live reference's training/metrics.py: identical floats."""
    d = load("metrics.npz")
    preds, tgts = unpad(d["pred"]), unpad(d["tgt"])
    M = oracle.metrics
    assert [M.levenshtein_distance(p, t) for p, t in zip(preds, tgts)] == d["lev"].tolist()
    assert [M.bleu_n_score(p, t, 4) for p, t in zip(preds, tgts)] == d["bleu4"].tolist()
    assert [M.bleu_n_score(p, t, 2) for p, t in zip(preds, tgts)] == d["bleu2"].tolist()
    res = M.calculate_metrics(preds[2:], tgts[2:])
    assert [res["bleu"], res["levenshtein"], float(res["batch_size"])] == d["mean"].tolist()


def test_validation_step_golden():
    """oracle.metrics.validation_loss_accuracy (+ oracle.seq2seq_forward) against one validation step of the live
    reference trainer's arithmetic (CrossEntropyLoss with label smoothing, masked_accuracy)."""
    d = load("validation.npz")
    cfg = H.SMALL
    formulas = torch.as_tensor(d["formulas"])
    p = oracle.make_params(cfg, 2, sharp=True)
    x = H.make_images(cfg, formulas.shape[0])
    check_inputs(d, p, x)
    with torch.no_grad():
        out = oracle.seq2seq_forward(p, x, formulas, cfg)
    close(out, d["outputs"])
    loss, correct, total = oracle.metrics.validation_loss_accuracy(torch.as_tensor(d["outputs"]), formulas[:, 1:], 0)
    assert abs(float(loss) - float(d["loss"])) < 1e-6 and (correct, total) == (int(d["correct"]), int(d["total"]))


def test_tokenizer_golden(pkg):
    """SURVEY 8f-2 (host side): the drop-in LaTeXTokenizer against the live reference's (data/tokenizer.py) default
    vocabulary, encode (with / without specials, unknown tokens -> UNK) and decode (specials skipped or kept,
    unknown ids -> the UNK string)."""
    d = load("tokenizer.npz")
    tok = pkg.LaTeXTokenizer()
    tok.default_init()
    assert [tok.id_to_token[i] for i in range(tok.vocab_size)] == d["vocab"].tolist()
    assert [tok.pad_token_id, tok.start_token_id, tok.end_token_id, tok.unk_token_id, tok.max_sequence_length] == d["specials"].tolist()
    texts = d["texts"].tolist()
    assert [tok.encode(t) for t in texts] == unpad(d["enc"])
    assert [tok.encode(t, add_special_tokens=True) for t in texts] == unpad(d["enc_sp"])
    ids = unpad(d["ids"])
    assert [tok.decode(i) for i in ids] == d["dec"].tolist()
    assert [tok.decode(i, skip_special_tokens=False) for i in ids] == d["dec_keep"].tolist()
    # from_config round trip (training/predictor.py:86-105)
    t2 = pkg.LaTeXTokenizer.from_config({"token_to_id": tok.token_to_id, "special_tokens": tok.special_tokens,
                                         "max_sequence_length": 77})
    assert t2.vocab_size == 46 and t2.max_sequence_length == 77 and t2.decode(ids[0]) == d["dec"][0]
