"""Device-side image preparation through the C-ABI (`i2l_resize_plan_build` +
`i2l_resize_pad_u8`) against the golden vectors of the LIVE reference (`ResizeWithAspectRatio`,
`load_image`, the PIL branch of `Predictor._prepare_image` -- all on live Pillow) and against the CPU
oracle (oracle/resize.py).  Integer / byte work: every comparison is bit-exact."""
import os

import numpy as np
import pytest
import torch

import helpers as H
from helpers import oracle

pytestmark = pytest.mark.gpu

G = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
R = oracle.resize


def golden_images(d):
    out, off = [], 0
    for h, w, c in d["shapes"]:
        n = int(h) * int(w) * int(c)
        a = d["pixels"][off:off + n]
        out.append(a.reshape(h, w, c) if c == 3 else a.reshape(h, w))
        off += n
    return out


def test_resize_with_aspect_ratio_vs_reference_golden(pkg):
    d = np.load(os.path.join(G, "resize.npz"))
    TH, TW = (int(v) for v in d["target"])
    imgs = golden_images(d)
    rwa = pkg.ResizeWithAspectRatio(TH, TW)
    gray = rwa([a for a in imgs if a.ndim == 2]).cpu().numpy()
    assert gray.shape == (9, 1, TH, TW) and np.array_equal(gray[:, 0], d["out_l"])
    rgb = rwa([a for a in imgs if a.ndim == 3]).cpu().numpy()
    assert np.array_equal(rgb.transpose(0, 2, 3, 1), d["out_rgb"])          # incl. the red RGB padding quirk


def test_load_images_vs_reference_golden(pkg):
    d = np.load(os.path.join(G, "resize.npz"))
    TH, TW = (int(v) for v in d["target"])
    files = [d["file0"], d["file1"]]
    one = pkg.load_images(files, (TH, TW), channels=1).cpu()                # convert("L") on the device
    three = pkg.load_images(files, (TH, TW), channels=3).cpu()              # ImageNet mean / std
    for j in range(2):
        assert torch.equal(one[j], torch.from_numpy(d[f"file{j}_load1"]))
        assert torch.equal(three[j], torch.from_numpy(d[f"file{j}_load3"]))


def test_predictor_prepare_image_kinds(pkg, tmp_path):
    """Predictor._prepare_image: PIL input (reference golden, 64x800 bicubic stretch), path input (load_image)."""
    from PIL import Image
    d = np.load(os.path.join(G, "resize.npz"))
    cfg = dict(model_type="cnn_lstm", vocab_size=46, embedding_dim=32, hidden_dim=48, lstm_layers=1, attention=True,
               img_height=64, img_width=800, channels=1, conv_filters=[4, 8, 8])
    m = H.build_model(pkg, cfg, oracle.make_params(cfg, 3))
    tok = pkg.LaTeXTokenizer(); tok.default_init()
    pred = pkg.Predictor(m, tok)
    got = pred._prepare_image(Image.fromarray(d["pil_in"], "RGB")).cpu()
    assert torch.equal(got, torch.from_numpy(d["pil_prepared"]))
    path = str(tmp_path / "f.png")
    Image.fromarray(d["file0"], "RGB").save(path)
    ref = oracle.normalize_u8(torch.from_numpy(R.resize_with_aspect_ratio(R.rgb_to_l(d["file0"]), 64, 800))[None, None], "pm1")
    assert torch.equal(pred._prepare_image(path).cpu(), ref)
    with pytest.raises(TypeError, match="Unsupported image type"):
        pred._prepare_image(3.5)
    strs = pred.predict_batch([path, Image.fromarray(d["pil_in"], "RGB"), torch.rand(1, 64, 800)], max_length=8)
    assert len(strs) == 3 and all(isinstance(s, str) for s in strs)


@pytest.mark.parametrize("channels,to_gray", [(1, False), (3, False), (3, True)])
@pytest.mark.parametrize("resample,mode", [("lanczos", "aspect"), ("bicubic", "stretch")])
def test_ragged_batch_vs_oracle(pkg, channels, to_gray, resample, mode):
    rng = np.random.default_rng(7 + channels)
    TH, TW = 64, 320
    imgs = []
    for i in range(24):
        h, w = int(rng.integers(1, 140)), int(rng.integers(3, 900))
        if mode == "aspect" and R.aspect_width(w, h, TH) < 1:
            w = h
        shape = (h, w, 3) if channels == 3 else (h, w)
        a = rng.integers(0, 256, size=shape, dtype=np.uint8) if i % 2 else (rng.random(shape) > 0.7).astype(np.uint8) * 255
        imgs.append(a)
    imgs.append(np.full((64, 320, 3) if channels == 3 else (64, 320), 17, np.uint8))      # both passes skipped
    imgs.append(np.zeros((0, 9, 3) if channels == 3 else (0, 9), np.uint8) if mode == "aspect" else imgs[0])  # h == 0
    P = pkg.preprocess
    out = P.ResizePlan(imgs, TH, TW, to_gray=to_gray, resample=resample, mode=mode).run("cuda").cpu().numpy()
    for a, got in zip(imgs, out):
        src = R.rgb_to_l(a) if to_gray else a
        if mode == "aspect":
            ref = R.resize_with_aspect_ratio(src, TH, TW)
        else:
            ref = R.resize_lanczos_u8(src, TW, TH, resample)
        ref = ref[None] if ref.ndim == 2 else ref.transpose(2, 0, 1)
        assert np.array_equal(got, ref), (a.shape, np.abs(got.astype(int) - ref.astype(int)).max())


def test_resize_properties_full_size(pkg):
    """Size-independent properties at the serving size (1024 images -> 64x800): constant images stay constant,
    outputs are idempotent under a second pass at the same size, and the batch result equals the
    image-by-image result (no cross-image leakage in the ragged launch)."""
    rng = np.random.default_rng(3)
    P = pkg.preprocess
    imgs = []
    for i in range(1024):
        h, w = int(rng.integers(20, 120)), int(rng.integers(60, 1400))
        if i % 4 == 0:
            imgs.append(np.full((h, w), i % 256, np.uint8))
        else:
            imgs.append(rng.integers(0, 256, size=(h, w), dtype=np.uint8))
    out = P.ResizePlan(imgs, 64, 800).run("cuda")
    torch.cuda.synchronize()
    o = out.cpu().numpy()
    for i in range(0, 1024, 4):           # the 22-bit weights sum to (1 << 22) +- a few units: v * sum rounds back to v
        nw = min(R.aspect_width(imgs[i].shape[1], imgs[i].shape[0], 64), 800)
        assert (o[i, 0, :, :nw] == i % 256).all()
        assert (o[i, 0, :, nw:] == 255).all()
    again = P.ResizePlan([o[i, 0] for i in range(64)], 64, 800).run("cuda").cpu().numpy()
    assert np.array_equal(again, o[:64])                          # same size: both passes are skipped
    for i in (1, 2, 3, 513, 1023):
        single = P.ResizePlan([imgs[i]], 64, 800).run("cuda").cpu().numpy()
        assert np.array_equal(single[0], o[i])
        assert np.array_equal(o[i, 0], R.resize_with_aspect_ratio(imgs[i], 64, 800))


def test_resize_errors(pkg):
    P = pkg.preprocess
    with pytest.raises(ValueError):
        P.ResizePlan([np.zeros((300, 1), np.uint8)], 16, 100)      # resized width 0: Pillow raises ValueError
    with pytest.raises(ValueError):
        P.ResizePlan([np.zeros((4, 4), np.uint8), np.zeros((4, 4, 3), np.uint8)], 16, 100)
    with pytest.raises(TypeError):
        P.ResizePlan([np.zeros((4, 4), np.float32)], 16, 100)
    assert P.ResizePlan([], 16, 100).run("cuda").shape == (0, 1, 16, 100)


def test_staging_pool_never_overwrites_an_unrun_plan(pkg):
    """three plans created before any of them runs must each keep their own pixels (the third one gets a
    private pinned buffer instead of recycling a slot an un-run plan still points into), in any run order."""
    import numpy as np
    P = pkg.preprocess
    g = np.random.default_rng(5)
    batches = [[g.integers(0, 256, (40 + 3 * k, 200 + 11 * k), dtype=np.uint8) for _ in range(3)] for k in range(4)]
    ref = [P.ResizePlan(b, 64, 320).run("cuda").cpu() for b in batches]
    plans = [P.ResizePlan(b, 64, 320) for b in batches]           # four un-run plans alive at once
    assert len({p.host.data_ptr() for p in plans}) == 4
    for k in (2, 0, 3, 1):
        assert torch.equal(plans[k].run("cuda").cpu(), ref[k])
