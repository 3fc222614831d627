"""Generates the golden vectors under tests/golden/ by executing the LIVE reference
(/root/reference, imported through oracle/ref_shim.py) on seeded inputs.  Run in the build
container only:  python tests/golden/make_golden.py

Inputs are not stored: they are regenerated from seeds by oracle.make_params /
helpers.make_images (a checksum of both is stored to catch generator drift).  Outputs of
the reference modules are stored as float32 / int64 arrays.
"""
import os
import sys

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(HERE))
import helpers as H  # noqa: E402
from helpers import oracle  # noqa: E402
from oracle import ref_shim  # noqa: E402

torch.set_num_threads(8)


def checksum(params, x):
    s = sum(float(v.double().abs().sum()) for v in params.values() if v.dtype.is_floating_point)
    return np.array([s, float(x.double().abs().sum())])


def save(name, **arrs):
    np.savez_compressed(os.path.join(HERE, name), **{k: (v.numpy() if isinstance(v, torch.Tensor) else np.asarray(v))
                                                     for k, v in arrs.items()})
    print("wrote", name, {k: np.asarray(v).shape for k, v in arrs.items()})


def pad_rows(rows, fill=-1):
    n = max(len(r) for r in rows)
    return np.array([list(r) + [fill] * (n - len(r)) for r in rows], dtype=np.int64)


@torch.no_grad()
def seq2seq_case(name, cfg, seed, sharp, B, T, beam=0, end_boost=0.0):
    p = oracle.make_params(cfg, seed, sharp=sharp)
    if end_boost:
        p["decoder.output_layer.bias"][H.END] += end_boost
    m = ref_shim.build_reference_model(cfg, p)
    x = H.make_images(cfg, B)
    enc = m.encoder(x)
    g = torch.Generator().manual_seed(5)
    tok = torch.randint(0, cfg["vocab_size"], (B, 1), generator=g)
    l1, (h1, c1) = m.decoder.decode_step(enc, tok, None)
    tok2 = torch.randint(0, cfg["vocab_size"], (B, 1), generator=g)
    l2, (h2, c2) = m.decoder.decode_step(enc, tok2, (h1, c1))
    raw = m.inference(x, H.START, H.END, max_length=T)                       # B>1: raw lists
    single = [m.inference(x[i:i + 1], H.START, H.END, max_length=T) for i in range(B)]   # B==1 post-processing
    arrs = dict(checksum=checksum(p, x), enc=enc, tok=tok, tok2=tok2, logits1=l1, h1=h1, c1=c1, logits2=l2, h2=h2,
                c2=c2, greedy_raw=pad_rows(raw), greedy_single=pad_rows(single), T=np.array(T),
                seed=np.array(seed), sharp=np.array(int(sharp)), B=np.array(B), end_boost=np.array(end_boost))
    if beam:
        bs = [m.inference(x[i:i + 1], H.START, H.END, max_length=T, beam_size=beam) for i in range(B)]
        arrs.update(beam=pad_rows(bs), beam_size=np.array(beam))
    save(name, **arrs)


@torch.no_grad()
def resnet_case(name, cfg, widths, B=2):
    p = oracle.make_params(cfg, 0)
    m = ref_shim.build_reference_model(cfg, p)
    arrs = {}
    for w in widths:
        x = H.make_images(cfg, B, width=w)
        arrs[f"enc_w{w}"] = m.encoder(x)
        arrs[f"checksum_w{w}"] = checksum(p, x)
    save(name, widths=np.array(widths), **arrs)


@torch.no_grad()
def attention_case(name):
    ref = ref_shim.load()
    g = torch.Generator().manual_seed(3)
    Hd, E, B, L = 48, 32, 6, 5
    att = ref["Attention"](Hd, E)
    w = oracle.port._uniform(g, (Hd, Hd + E), 0.2); b = oracle.port._uniform(g, (Hd,), 0.2)
    v = oracle.port._uniform(g, (1, Hd), 0.3)
    att.load_state_dict({"attn.weight": w, "attn.bias": b, "v.weight": v})
    hid = torch.randn(B, 1, Hd, generator=g); enc = torch.randn(B, L, E, generator=g)
    save(name, w=w, b=b, v=v, hid=hid, enc=enc, ctx=att(hid, enc))


@torch.no_grad()
def predict_batch_case(name, temperature, top_k, top_p, T=16, B=5):
    """Runs the reference's Predictor.predict_batch (predictor.py:205-394) with
    torch.multinomial replaced by the restated inverse-CDF draw on a fixed uniform stream,
    recording the filtered distribution it is handed at every step."""
    ref = ref_shim.load()
    cfg = dict(model_type="cnn_lstm", vocab_size=46, embedding_dim=32, hidden_dim=48, lstm_layers=1, attention=True,
               img_height=64, img_width=800, channels=1, conv_filters=[4, 8, 8])
    p = oracle.make_params(cfg, 3, sharp=True)
    m = ref_shim.build_reference_model(cfg, p)
    tok = ref["LaTeXTokenizer"](); tok.default_init()
    pred = ref["Predictor"](m, tok, device=torch.device("cpu"), model_type="cnn_lstm")
    g = torch.Generator().manual_seed(11)
    imgs = [torch.rand(1, 64, 800, generator=g) for _ in range(B)]             # in [0,1]: passes _preprocess_tensor untouched
    u = torch.rand(T, B, generator=torch.Generator().manual_seed(9))
    rec, step = [], [0]
    real_multinomial = torch.multinomial

    def fake_multinomial(probs, n):
        rec.append(probs.clone())
        out = oracle.inverse_cdf_draw(probs, u[step[0]]).unsqueeze(1)
        step[0] += 1
        return out

    torch.multinomial = fake_multinomial
    try:
        strs = pred.predict_batch(imgs, max_length=T, temperature=temperature, top_k=top_k, top_p=top_p, batch_size=B)
    finally:
        torch.multinomial = real_multinomial
    arrs = dict(strings=np.array(strs), u=u, temperature=np.array(temperature), top_k=np.array(top_k),
                top_p=np.array(top_p), T=np.array(T), B=np.array(B))
    if rec:
        arrs["probs"] = torch.stack(rec)
    save(name, **arrs)


@torch.no_grad()
def teacher_forced_case(name):
    """Reference `Seq2SeqModel.forward` in eval mode (seq2seq.py:98-122 -> decoder.py:100-195): both decoder
    branches (per-step loop with attention; one nn.LSTM call over the sequence without)."""
    arrs = {}
    for tag, cfg, seed, B, T in (("headline", H.HEADLINE, 1, 3, 12), ("small_l2", H.SMALL, 2, 5, 9),
                                 ("small_l2_noattn", dict(H.SMALL, attention=False), 2, 5, 9)):
        p = oracle.make_params(cfg, seed, sharp=True)
        m = ref_shim.build_reference_model(cfg, p)
        x = H.make_images(cfg, B)
        tgt = torch.randint(0, cfg["vocab_size"], (B, T + 1), generator=torch.Generator().manual_seed(21))
        arrs[f"{tag}_checksum"] = checksum(p, x)
        arrs[f"{tag}_target"] = tgt
        arrs[f"{tag}_logits"] = m(x, tgt)
        arrs[f"{tag}_meta"] = np.array([seed, B, T])
    save(name, **arrs)


@torch.no_grad()
def validation_case(name):
    """One validation step of the reference trainer (training/trainer.py:517-529): model(images, formulas) ->
    CrossEntropyLoss(ignore_index=pad, reduction="mean", label_smoothing=0.1) (trainer.py:111-115) and
    masked_accuracy (training/metrics.py:226-238)."""
    masked_accuracy = ref_shim.reference_module("img2latex.training.metrics").masked_accuracy
    cfg, B, T = H.SMALL, 6, 12
    p = oracle.make_params(cfg, 2, sharp=True)
    m = ref_shim.build_reference_model(cfg, p)
    x = H.make_images(cfg, B)
    g = torch.Generator().manual_seed(33)
    formulas = torch.randint(4, cfg["vocab_size"], (B, T + 1), generator=g)
    formulas[:, 0] = H.START
    for b in range(B):
        e = int(torch.randint(3, T + 1, (1,), generator=g))
        formulas[b, e] = H.END
        formulas[b, e + 1:] = 0                                   # PAD
    outputs = m(x, formulas)
    targets = formulas[:, 1:]
    crit = torch.nn.CrossEntropyLoss(ignore_index=0, reduction="mean", label_smoothing=0.1)
    loss = crit(outputs.transpose(1, 2), targets)
    correct, total = masked_accuracy(outputs, targets, 0)
    save(name, formulas=formulas, outputs=outputs, loss=np.array(float(loss)), correct=np.array(correct), total=np.array(total),
         checksum=checksum(p, x))


def load_image_case(name):
    """Reference `load_image` (data/utils.py:18-90) on PNG files whose size already equals the
    target size (ResizeWithAspectRatio is then the identity, transforms.py:38-43)."""
    import tempfile
    from PIL import Image
    load_image = ref_shim.reference_module("img2latex.data.utils").load_image
    g = np.random.default_rng(4)
    rgb = g.integers(0, 256, size=(16, 48, 3), dtype=np.uint8)
    gray = g.integers(0, 256, size=(16, 48), dtype=np.uint8)
    with tempfile.TemporaryDirectory() as td:
        Image.fromarray(rgb, "RGB").save(os.path.join(td, "rgb.png"))
        Image.fromarray(gray, "L").save(os.path.join(td, "gray.png"))
        out_rgb = load_image(os.path.join(td, "rgb.png"), (16, 48), 3)
        out_gray = load_image(os.path.join(td, "gray.png"), (16, 48), 1)
    save(name, rgb=rgb, gray=gray, out_rgb=out_rgb, out_gray=out_gray)


def formula_like(rng, h, w, c):
    """White page with dark strokes and a little grey (compressible, but exercises ringing / clipping)."""
    a = np.full((h, w, c), 255, np.uint8)
    for _ in range(max(1, (h * w) // 150)):
        y, x = int(rng.integers(0, h)), int(rng.integers(0, w))
        dh, dw = int(rng.integers(1, max(2, h // 4))), int(rng.integers(1, max(2, w // 8)))
        a[y:y + dh, x:x + dw] = rng.integers(0, 120, size=c if rng.random() < 0.3 else 1)
    return a if c == 3 else a[:, :, 0]


def resize_case(name):
    """Reference `ResizeWithAspectRatio` (data/transforms.py:9-56; live Pillow), `load_image`
    (data/utils.py:18-90, through PNG files) and the PIL branch of `Predictor._prepare_image`
    (training/predictor.py:427-446) on seeded ragged images."""
    import tempfile
    from PIL import Image
    RWA = ref_shim.reference_module("img2latex.data.transforms").ResizeWithAspectRatio
    load_image = ref_shim.reference_module("img2latex.data.utils").load_image
    rng = np.random.default_rng(12)
    arrs = {}
    # (h, w, channels): pad, crop, identity width, identity height, up- and down-scaling, 1-pixel edges
    sizes = [(20, 90, 1), (50, 400, 1), (64, 300, 1), (33, 700, 1), (100, 330, 1), (7, 9, 1), (1, 40, 1), (40, 1, 1),
             (16, 100, 1), (30, 160, 3), (80, 200, 3), (12, 300, 3), (64, 64, 3), (45, 211, 3)]
    TH, TW = 32, 160
    shapes, flat, outs_l, outs_rgb = [], [], [], []
    for i, (h, w, c) in enumerate(sizes):
        a = formula_like(rng, h, w, c) if i % 3 else rng.integers(0, 256, size=(h, w, c) if c == 3 else (h, w), dtype=np.uint8)
        shapes.append((h, w, c)); flat.append(a.reshape(-1))
        img = Image.fromarray(a, "L" if c == 1 else "RGB")
        out = np.array(RWA(TH, TW)(img))
        (outs_l if c == 1 else outs_rgb).append(out)
    arrs.update(shapes=np.array(shapes), pixels=np.concatenate(flat), target=np.array([TH, TW]),
                out_l=np.stack(outs_l), out_rgb=np.stack(outs_rgb))
    # load_image through files: RGB file read with channels=1 (convert L on the way) and channels=3 (ImageNet norm)
    with tempfile.TemporaryDirectory() as td:
        li = []
        for j, (h, w) in enumerate([(24, 150), (50, 90)]):
            a = formula_like(rng, h, w, 3)
            Image.fromarray(a, "RGB").save(os.path.join(td, f"f{j}.png"))
            arrs[f"file{j}"] = a
            arrs[f"file{j}_load1"] = load_image(os.path.join(td, f"f{j}.png"), (TH, TW), 1)
            arrs[f"file{j}_load3"] = load_image(os.path.join(td, f"f{j}.png"), (TH, TW), 3)
    # PIL branch of _prepare_image: plain resize with Pillow's default filter (bicubic), /255*2-1
    ref = ref_shim.load()
    cfg = dict(model_type="cnn_lstm", vocab_size=46, embedding_dim=32, hidden_dim=48, lstm_layers=1, attention=True,
               img_height=64, img_width=800, channels=1, conv_filters=[4, 8, 8])
    m = ref_shim.build_reference_model(cfg, oracle.make_params(cfg, 3))
    tok = ref["LaTeXTokenizer"](); tok.default_init()
    pred = ref["Predictor"](m, tok, device=torch.device("cpu"), model_type="cnn_lstm")
    a = formula_like(rng, 40, 230, 3)
    arrs["pil_in"] = a
    arrs["pil_prepared"] = pred._prepare_image(Image.fromarray(a, "RGB"))        # (1,1,64,800) fp32
    save(name, **arrs)


def checkpoint_case(name):
    """Reference `Predictor.from_checkpoint` + `predict` / `predict_batch` (training/predictor.py:61-203) on a
    checkpoint file in the reference's own layout."""
    import tempfile
    ref = ref_shim.load()
    p = oracle.make_params(H.CKPT_CFG, 2, sharp=True)
    g = torch.Generator().manual_seed(13)
    imgs = [torch.rand(1, 64, 800, generator=g) for _ in range(4)]
    with tempfile.TemporaryDirectory() as td:
        path = os.path.join(td, "best_checkpoint_epoch_3_step_7.pt")
        H.make_checkpoint(path, p)
        pred = ref["Predictor"].from_checkpoint(path, device=torch.device("cpu"))
        singles = []
        for im in imgs:
            try:
                singles.append(pred.predict(im, max_length=24))
            except IndexError:        # predictor.py:194 indexes an EMPTY sequence when the first token is END
                singles.append("<IndexError>")
        batch = pred.predict_batch(imgs, max_length=24, batch_size=4)
    save(name, singles=np.array(singles), batch=np.array(batch), vocab_size=np.array(pred.tokenizer.vocab_size))


def tokenizer_case(name):
    """Reference `LaTeXTokenizer` (data/tokenizer.py): default vocabulary, encode / decode round trips, special-token
    handling -- the detokenise half of SURVEY 8f-2."""
    ref = ref_shim.load()
    tok = ref["LaTeXTokenizer"]()
    tok.default_init()
    vocab = [tok.id_to_token[i] for i in range(tok.vocab_size)]
    texts = ["\\frac { a } { b } + \\sum _ { x = 0 } ^ { \\infty } x", "x ^ 2 + unknown_token - y", "", "<START> a <END> <PAD>"]
    enc = [tok.encode(t) for t in texts]
    enc_sp = [tok.encode(t, add_special_tokens=True) for t in texts]
    ids = [[1, 16, 4, 19, 2, 0, 0], [3, 45, 44, 7, 2], [], [1, 2], [99, 5]]
    dec = [tok.decode(i) for i in ids]
    dec_keep = [tok.decode(i, skip_special_tokens=False) for i in ids]
    save(name, vocab=np.array(vocab), texts=np.array(texts), enc=pad_rows(enc + [[0]])[:-1], enc_sp=pad_rows(enc_sp + [[0]])[:-1],
         ids=pad_rows(ids + [[0]])[:-1], dec=np.array(dec), dec_keep=np.array(dec_keep),
         specials=np.array([tok.pad_token_id, tok.start_token_id, tok.end_token_id, tok.unk_token_id, tok.max_sequence_length]))


def metrics_pairs(seed=31, n=48):
    """Seeded (prediction, target) id lists: small vocabularies (many repeated n-grams), empty / one-token /
    identical / prefix / long cases."""
    rng = np.random.default_rng(seed)
    pairs = [([], []), ([], [5, 6]), ([5], []), ([7], [7]), ([3, 3, 3, 3, 3], [3, 3, 3]), ([1, 2, 3, 4], [1, 2, 3, 4]),
             ([1, 2, 3], [1, 2, 3, 4, 5, 6]), ([9, 8, 7, 6, 5, 4], [4, 5, 6, 7, 8, 9])]
    while len(pairs) < n:
        V = int(rng.choice([2, 3, 6, 40]))
        t = rng.integers(0, V, size=int(rng.integers(1, 60))).tolist()
        kind = rng.random()
        if kind < 0.5:                      # noisy copy of the target
            p = [x for x in t if rng.random() > 0.15]
            p = [int(rng.integers(0, V)) if rng.random() < 0.15 else x for x in p]
        else:
            p = rng.integers(0, V, size=int(rng.integers(1, 60))).tolist()
        pairs.append((p, t))
    pairs.append((rng.integers(0, 5, size=300).tolist(), rng.integers(0, 5, size=270).tolist()))   # > 256 columns
    return pairs


def metrics_case(name):
    """Reference training/metrics.py (levenshtein_distance 49-94, bleu_n_score 97-179, calculate_metrics 182-223)."""
    M = ref_shim.reference_module("img2latex.training.metrics")
    pairs = metrics_pairs()
    lev = np.array([M.levenshtein_distance(p, t) for p, t in pairs], np.float64)
    bleu4 = np.array([M.bleu_n_score(p, t, 4) for p, t in pairs], np.float64)
    bleu2 = np.array([M.bleu_n_score(p, t, 2) for p, t in pairs], np.float64)
    res = M.calculate_metrics([p for p, _ in pairs[2:]], [t for _, t in pairs[2:]])
    save(name, pred=pad_rows([p for p, _ in pairs] + [[0]])[:-1], tgt=pad_rows([t for _, t in pairs] + [[0]])[:-1],
         lev=lev, bleu4=bleu4, bleu2=bleu2, mean=np.array([res["bleu"], res["levenshtein"], res["batch_size"]], np.float64))


if __name__ == "__main__":
    assert ref_shim.available(), "the live reference is needed to (re)generate golden vectors"
    seq2seq_case("cnn_headline_sharp.npz", H.HEADLINE, seed=1, sharp=True, B=4, T=30)
    seq2seq_case("cnn_headline_default.npz", H.HEADLINE, seed=0, sharp=False, B=3, T=12)
    seq2seq_case("cnn_small_l2_beam.npz", H.SMALL, seed=2, sharp=True, B=6, T=20, beam=3, end_boost=0.0)
    resnet_case("resnet18.npz", H.R18, [128, 160])
    resnet_case("resnet50.npz", H.R50, [96])
    attention_case("attention_L5.npz")
    load_image_case("load_image.npz")
    teacher_forced_case("teacher_forced.npz")
    resize_case("resize.npz")
    metrics_case("metrics.npz")
    checkpoint_case("checkpoint.npz")
    validation_case("validation.npz")
    tokenizer_case("tokenizer.npz")
    predict_batch_case("predict_batch_greedy.npz", 1.0, 0, 0.0)
    predict_batch_case("predict_batch_topk_topp.npz", 0.8, 5, 0.9)
    predict_batch_case("predict_batch_topp.npz", 1.2, 0, 0.7)
