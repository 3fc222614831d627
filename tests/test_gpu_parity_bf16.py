"""bf16 fast path (tcgen05 kernels) against the fp32 CPU oracle.

Stated bf16 tolerance: encoder outputs within 3e-2 of the row max-abs; greedy
token rows may leave the fp32 oracle's sequence only at a step where the oracle's top-1/top-2 logit
margin is below BF16_MARGIN * max|logit| (near tie under bf16 rounding) -- such rows are
counted and excluded from there on; every other row must match token for token.

Two oracles are used.  (1) the fp32 restatement of the reference: a random-init decoder has many
near ties (F13), and one flipped token changes the rest of a row, so the fraction of rows that stay
on the fp32 sequence for a whole decode is a property of the weights, not of the kernel; what the
kernel must guarantee is that EVERY departure sits at a near tie.  (2) the same restatement with
the GEMM operands rounded to bf16 where the kernels round them (oracle `operand_rounding="bf16"`):
this removes the rounding that is inherent to the mode, so against it the kernels must reproduce
>= 90 % of the rows token for token (what remains is accumulation order and MUFU.TANH)."""
import pytest
import torch

import helpers as H
from helpers import oracle

pytestmark = pytest.mark.gpu

BF16_MARGIN = 2.5e-2      # measured: the largest oracle margin at which a row left the fp32 sequence is 2.4e-3 in the
                          # short cases and 1.7e-2 over 1024 rows x 150 steps (the cell state integrates the rounding)


def bf16_cfg(cfg):
    return dict(cfg, operand_rounding="bf16")


def rows_identical(tokens, ref_seqs):
    return sum(tokens[b][: len(r)] == r for b, r in enumerate(ref_seqs))


def divergence_report(tokens, ref_seqs, logit_trace):
    exact, near, bad = 0, [], []
    for b, ref in enumerate(ref_seqs):
        got = tokens[b][: len(ref)]
        if got == ref:
            exact += 1
            continue
        t = next(i for i, (a, r) in enumerate(zip(got, ref)) if a != r) - 1
        lg = logit_trace[t][b]
        top2 = torch.topk(lg, 2).values
        rel = float(top2[0] - top2[1]) / float(lg.abs().max())
        (near if rel < BF16_MARGIN else bad).append((b, t, round(rel, 5)))
    return exact, near, bad


@pytest.mark.parametrize("B,T,seed,sharp", [(32, 40, 1, True), (50, 60, 2, True), (128, 30, 0, False), (7, 25, 1, True)])
def test_persistent_greedy_vs_oracle(pkg, B, T, seed, sharp):
    cfg = H.HEADLINE
    p = oracle.make_params(cfg, seed, sharp=sharp)
    m32 = H.build_model(pkg, cfg, p, precision="fp32")
    m16 = H.build_model(pkg, cfg, p, precision="bf16")
    x = H.make_images(cfg, B)
    enc_ref = oracle.encoder(p, x, cfg)
    ref, steps_ref, trace = oracle.greedy_search(p, enc_ref, H.START, H.END, T, 1.0, cfg, return_logits=True)
    enc = m32.encoder(x.cuda())                       # isolate the decoder: fp32 encodings in
    tokens, lengths, steps = m16.decoder.greedy(enc, H.START, H.END, T)
    torch.cuda.synchronize()
    toks = tokens.cpu()[:, : steps_ref + 1].tolist()
    exact, near, bad = divergence_report(toks, ref, trace)
    ref16, steps16 = oracle.greedy_search(p, enc_ref, H.START, H.END, T, 1.0, bf16_cfg(cfg))
    same16 = rows_identical(tokens.cpu()[:, : steps16 + 1].tolist(), ref16)
    print(f"bf16 persistent greedy B={B} T={T}: rows on the fp32 oracle {exact}/{B}, near-tie divergences {near}; "
          f"rows on the bf16-operand oracle {same16}/{B}")
    assert not bad, f"rows diverging at a step with a clear margin: {bad}"
    assert exact >= 0.7 * B                           # measured 0.71-0.84 (random-init near ties, see the module docstring)
    assert same16 >= 0.9 * B
    if not near:
        assert int(steps) == steps_ref
        for b in range(B):
            exp = ref[b].index(H.END) if H.END in ref[b][1:] else steps_ref + 1
            assert int(lengths[b]) == exp


def test_persistent_sticky_stop_and_lengths(pkg):
    cfg = H.HEADLINE
    p = oracle.make_params(cfg, 2, sharp=True)
    p["decoder.output_layer.bias"][H.END] += 2.0      # every row emits END early: the sticky rule fires
    m16 = H.build_model(pkg, cfg, p, precision="bf16")
    m32 = H.build_model(pkg, cfg, p, precision="fp32")
    B, T = 70, 60
    x = H.make_images(cfg, B)
    enc = m32.encoder(x.cuda())
    N = pkg._native
    t16, l16, s16 = m16.decoder.greedy(enc, H.START, H.END, T, 1.0, N.STOP_ALL_FINISHED_STICKY)
    t32, l32, s32 = m32.decoder.greedy(enc, H.START, H.END, T, 1.0, N.STOP_ALL_FINISHED_STICKY)
    assert int(s32) < T, "test premise: the sticky rule fires early"
    agree = 0
    for b in range(B):
        a, r = t16[b, : int(l16[b])].tolist(), t32[b, : int(l32[b])].tolist()
        agree += a == r
    assert agree >= 0.9 * B
    if agree == B:
        assert int(s16) == int(s32)


@pytest.mark.parametrize("B,sharp", [(2, False), (1, True), (5, True), (16, True), (33, False)])
def test_cnn_encoder_bf16(pkg, B, sharp):
    """tcgen05 implicit-GEMM conv stack + split-K FC vs the fp32 oracle: within 3e-2 of max|out|."""
    cfg = H.HEADLINE
    p = oracle.make_params(cfg, 1, sharp=sharp)
    m16 = H.build_model(pkg, cfg, p, precision="bf16")
    x = H.make_images(cfg, B)
    ref = oracle.cnn_encoder(p, x)
    out = m16.encoder(x.cuda())
    torch.cuda.synchronize()
    err = H.rel_err(out, ref)
    print(f"bf16 encoder B={B}: rel err {err:.3e}")
    assert out.shape == ref.shape and err < 3e-2


@pytest.mark.parametrize("B", [1, 6, 33])
def test_cnn_encoder_bf16_input_matches_fp32_input(pkg, B):
    """bf16 image tensors (half the H2D / HBM bytes) feed conv1 directly; since the conv1
    operand is bf16 either way the result must be BIT-identical to feeding the same values
    as fp32, and within the bf16 tolerance of the fp32 oracle on the unrounded images."""
    cfg = H.HEADLINE
    p = oracle.make_params(cfg, 1, sharp=True)
    m16 = H.build_model(pkg, cfg, p, precision="bf16")
    x = H.make_images(cfg, B)
    xb = x.bfloat16()
    out_b = m16.encoder(xb.cuda())
    out_f = m16.encoder(xb.float().cuda())
    torch.cuda.synchronize()
    assert torch.equal(out_b, out_f)
    assert H.rel_err(out_b, oracle.cnn_encoder(p, x)) < 3e-2


@pytest.mark.parametrize("shape,mode,nhwc", [((3, 3, 64, 320), "pm1", False), ((2, 3, 64, 320), "meanstd", False),
                                             ((2, 1, 17, 23), "pm1", False), ((3, 3, 9, 31), "meanstd", True),
                                             ((2, 3, 64, 320), "meanstd", True)])
def test_normalize_u8(pkg, shape, mode, nhwc):
    """i2l_normalize_u8 vs the reference's load_image arithmetic (oracle.normalize_u8, pinned by
    tests/golden/load_image.npz): fp32 output bit-exact, bf16 output = RN(fp32 result)."""
    g = torch.Generator().manual_seed(3)
    px = torch.randint(0, 256, shape, dtype=torch.uint8, generator=g)
    ref = oracle.normalize_u8(px, mode)
    src = px.permute(0, 2, 3, 1).contiguous() if nhwc else px
    out32 = pkg.normalize_u8(src.cuda(), mode, channels_last=nhwc)
    out16 = pkg.normalize_u8(src.cuda(), mode, channels_last=nhwc, out_dtype=torch.bfloat16)
    torch.cuda.synchronize()
    assert torch.equal(out32.cpu(), ref)
    assert torch.equal(out16.cpu(), ref.bfloat16())


def _check_fused_u8(pkg, cfg, p, px, mode):
    """forward_u8 runs conv1 with an FP16 operand built from the raw bytes (cnn_bf16.cu: conv1_u8_kernel), the two-step
    path with the bf16 tensor.  (1) Both within the stated bf16 tolerance (3e-2 of max|out|) of the fp32 oracle, the
    fused one no worse than the two-step one; (2) the two differ by operand rounding only (measured ~3e-3), a
    border / padding bug moves whole output rows; (3) with a normalisation and conv1 weights that are exact in bf16
    AND fp16 (x / 128, weights on a 2^-8 grid) the two paths multiply identical numbers and must agree bit for bit,
    including the zero padding in NORMALISED space at the image border (raw 0 is not padding)."""
    m16 = H.build_model(pkg, cfg, p, precision="bf16")
    ref = oracle.cnn_encoder(p, oracle.normalize_u8(px, mode))
    fused = m16.encoder.forward_u8(px.cuda(), mode)
    two_step = m16.encoder(pkg.normalize_u8(px.cuda(), mode, out_dtype=torch.bfloat16))
    torch.cuda.synchronize()
    e_f, e_t, d = H.rel_err(fused, ref), H.rel_err(two_step, ref), H.rel_err(fused, two_step)
    print(f"fused-u8 vs oracle {e_f:.2e}, two-step bf16 vs oracle {e_t:.2e}, fused vs two-step {d:.2e}")
    assert e_f < 3e-2 and e_f < 1.5 * e_t + 1e-3 and d < 1e-2
    # exact-operand case
    q = {k: v.clone() for k, v in p.items()}
    wkey = "encoder.cnn_layers.0.weight"
    q[wkey] = (q[wkey].clamp(-0.99, 0.99) * 256).round() / 256
    mq = H.build_model(pkg, cfg, q, precision="bf16")
    mean, std = (0.0, 0.0, 0.0), (128.0 / 255.0,) * 3
    fused = mq.encoder.forward_u8(px.cuda(), "meanstd", mean, std)
    two_step = mq.encoder(pkg.normalize_u8(px.cuda(), "meanstd", mean, std, out_dtype=torch.bfloat16))
    torch.cuda.synchronize()
    assert torch.equal(fused, two_step)


@pytest.mark.parametrize("B,mode", [(5, "pm1"), (4, "meanstd"), (1, "pm1")])
def test_cnn_fused_uint8_conv1(pkg, B, mode):
    """CNNEncoder.forward_u8 (normalisation fused into the tcgen05 conv1, i2l_cnn_encoder_fwd_u8) at the headline shape."""
    cfg = H.HEADLINE
    p = oracle.make_params(cfg, 1)
    g = torch.Generator().manual_seed(11 + B)
    px = torch.randint(0, 256, (B, 3, 64, 320), dtype=torch.uint8, generator=g)
    px[:, :, 0, :] = 0; px[:, :, -1, :] = 255; px[:, :, :, 0] = 0; px[:, :, :, -1] = 7      # borders: raw 0 is not padding
    _check_fused_u8(pkg, cfg, p, px, mode)


def test_greedy_stream_uint8_and_bf16_hosts(pkg):
    """Seq2SeqModel.greedy_stream with raw uint8 pixels / bf16 tensors on the host gives the same
    tokens as the fp32-tensor call on the normalised images."""
    cfg = H.HEADLINE
    p = oracle.make_params(cfg, 1, sharp=True)
    m16 = H.build_model(pkg, cfg, p, precision="bf16")
    g = torch.Generator().manual_seed(8)
    px = torch.randint(0, 256, (9, 3, 64, 320), dtype=torch.uint8, generator=g)
    xn = oracle.normalize_u8(px, "pm1").bfloat16()
    enc = m16.encoder(xn.float().cuda())
    t_ref, l_ref, s_ref = m16.decoder.greedy(enc, H.START, H.END, 20)
    for host in (px.pin_memory(), xn.pin_memory()):
        (tok, lens, steps), = list(m16.greedy_stream([host], H.START, H.END, 20))
        assert torch.equal(tok, t_ref.cpu()) and torch.equal(lens, l_ref.cpu()) and steps == int(s_ref)


# ---- bf16 tcgen05 ResNet trunk (resnet_bf16.cu): im2col-TMA implicit GEMM, NHWC bf16 activations.
# Stated tolerance: encoder output within 3e-2 of max|out| of the fp32 oracle (measured 3e-3 for
# resnet18, 6e-3 for resnet50: every activation is rounded to bf16 once per layer).
RESNET_BF16_TOL = 3e-2


@pytest.mark.parametrize("cfg,B,width", [(H.R18, 3, 128), (H.R18, 5, 224), (H.R18, 2, 320), (H.R18, 37, 160),
                                         (H.R50, 2, 96), (H.R50, 9, 192)])
def test_resnet_bf16_encoder(pkg, cfg, B, width):
    p = oracle.make_params(cfg, 0)
    m = H.build_model(pkg, cfg, p, precision="bf16")
    x = H.make_images(cfg, B, width=width)
    ref = oracle.resnet_encoder(p, x, cfg["model_name"])
    lib = pkg._native.lib()
    l0 = lib.i2l_launch_count()
    out = m.encoder(x.cuda())
    torch.cuda.synchronize()
    err = H.rel_err(out, ref)
    print(f"{cfg['model_name']} bf16 B={B} W={width}: rel err {err:.3e}, launches {lib.i2l_launch_count() - l0}")
    assert out.shape == ref.shape
    assert err < RESNET_BF16_TOL


def test_resnet_bf16_rows_are_batch_independent(pkg):
    """An image's encoding must not depend on its batch neighbours (im2col tiles run across image
    boundaries): bit-identical rows for the same image at different batch positions / sizes."""
    cfg = H.R18
    p = oracle.make_params(cfg, 1)
    m = H.build_model(pkg, cfg, p, precision="bf16")
    x = H.make_images(cfg, 6, width=192).cuda()
    full = m.encoder(x)
    part = m.encoder(x[2:5].contiguous())
    assert torch.equal(full[2:5], part)
    assert torch.equal(m.encoder(x[5:6].contiguous()), full[5:6])


def test_resnet_bf16_odd_width_uses_fp32_path(pkg):
    """The pixel-pair stem needs an even width; other widths run on the fp32 kernels (still CUDA)."""
    cfg = H.R18
    p = oracle.make_params(cfg, 0)
    m = H.build_model(pkg, cfg, p, precision="bf16")
    x = H.make_images(cfg, 2, width=131)
    ref = oracle.resnet_encoder(p, x, cfg["model_name"])
    assert H.rel_err(m.encoder(x.cuda()), ref) < 1e-3


def test_resnet_forward_buckets_matches_forward(pkg):
    """Width buckets on several streams (ResNetEncoder.forward_buckets) give the same bits as one forward per bucket."""
    cfg = H.R18
    p = oracle.make_params(cfg, 0)
    m = H.build_model(pkg, cfg, p, precision="bf16")
    xs = [H.make_images(cfg, b, seed=w, width=w).cuda() for b, w in ((3, 128), (5, 160), (2, 320), (4, 224), (1, 96), (6, 192))]
    ref = [m.encoder(x).clone() for x in xs]
    for use_graphs in (False, True, True):          # second graph pass = pure replays
        outs = m.encoder.forward_buckets(xs, n_streams=3, use_graphs=use_graphs)
        torch.cuda.synchronize()
        for a, b in zip(outs, ref):
            assert torch.equal(a, b)
    # replays pick up new data in the static inputs
    xs2 = [x.flip(0).contiguous() for x in xs]
    outs = m.encoder.forward_buckets(xs2, n_streams=3)
    torch.cuda.synchronize()
    for a, b in zip(outs, ref):
        assert torch.equal(a, b.flip(0))


# ---- general decode loops with the per-step GEMMs on tcgen05 (gemm_bf16.cu): any E / H / L, precision "bf16"
@pytest.mark.parametrize("cfg,B,T", [(H.SMALL, 12, 20), (H.R18, 9, 16), (H.HEADLINE, 140, 12)])
def test_general_bf16_sampling_and_greedy(pkg, cfg, B, T):
    """Sampling (always the general path) and, for shapes the persistent kernel does not cover, greedy: bf16 weights
    and h on the tensor cores, fp32 accumulation / cell state / epilogue.  Filtered distributions within 6e-2 of the
    oracle's per-row maximum up to the first divergence of a row (the bf16 rounding of h feeds back through the
    recurrence: measured 3e-2 after 15 steps with the sharpened output layer); greedy rows may leave the oracle only
    at near ties."""
    p = oracle.make_params(cfg, 1, sharp=True)
    m32 = H.build_model(pkg, cfg, p, precision="fp32")
    m16 = H.build_model(pkg, cfg, p, precision="bf16")
    x = H.make_images(cfg, min(B, 16))
    enc_ref = oracle.encoder(p, x, cfg).repeat((B + 15) // 16, 1)[:B]
    enc = m32.encoder(x.cuda()).repeat((B + 15) // 16, 1)[:B].contiguous()
    u = torch.rand(T, B, generator=torch.Generator().manual_seed(4))
    seqs, trimmed, steps_ref, ptrace = oracle.sample_loop(p, enc_ref, H.START, H.END, T, 0.9, 20, 0.9, cfg, uniforms=u,
                                                          return_probs=True)
    tokens, lengths, steps, probs = m16.decoder.sample(enc, H.START, H.END, T, 0.9, 20, 0.9, uniforms=u, return_probs=True)
    tokens, probs = tokens.cpu(), probs.cpu()
    same_rows = 0
    for b in range(B):
        ref_row = seqs[b].tolist()
        got_row = tokens[b, : len(ref_row)].tolist()
        t_div = next((i for i, (a, r) in enumerate(zip(got_row, ref_row)) if a != r), None)
        same_rows += t_div is None
        upto = min(len(ref_row) - 1 if t_div is None else t_div, steps_ref)
        for t in range(upto):
            # the kept SET may differ by entries at the top-k / top-p cut (bf16 logits): compare where both keep
            both = (probs[t, b] > 0) & (ptrace[t][b] > 0)
            assert both.any()
            d = (probs[t, b] - ptrace[t][b]).abs()[both].max() / ptrace[t][b].max()
            assert float(d) < 6e-2, (b, t, float(d))
    seqs16, _, _ = oracle.sample_loop(p, enc_ref, H.START, H.END, T, 0.9, 20, 0.9, bf16_cfg(cfg), uniforms=u)
    same16 = sum(tokens[b, : seqs16.shape[1]].tolist() == seqs16[b].tolist() for b in range(B))
    print(f"general bf16 sampling {cfg.get('model_name', 'cnn')} B={B}: {same_rows}/{B} rows follow the fp32 oracle, "
          f"{same16}/{B} the bf16-operand oracle")
    # measured: 10/12, 8/9, 116/140 on the fp32 oracle; 8/12, 8/9, 140/140 on the bf16-operand one (the stream-ordered
    # path keeps the context term in fp32, so for the tiny configs neither oracle rounds exactly like it)
    assert max(same_rows, same16) >= 0.8 * B, f"only {same_rows} / {same16} of {B} sampled rows follow the oracles"
    if not pkg._native.lib().i2l_device_check() and cfg is not H.HEADLINE:     # greedy takes the general path too
        ref, steps_ref, trace = oracle.greedy_search(p, enc_ref, H.START, H.END, T, 1.0, cfg, return_logits=True)
        tok, _, _ = m16.decoder.greedy(enc, H.START, H.END, T)
        exact, near, bad = divergence_report(tok.cpu()[:, : steps_ref + 1].tolist(), ref, trace)
        assert not bad, f"rows diverging at a step with a clear margin: {bad}"
        assert exact >= 0.7 * B


@pytest.mark.parametrize("B,T,temperature,top_k,top_p", [(40, 30, 0.9, 20, 0.9), (64, 40, 0.8, 50, 0.9), (33, 25, 1.0, 0, 0.7),
                                                         (32, 20, 1.3, 5, 0.0), (16, 20, 0.7, 0, 0.0), (70, 24, 1.0, 100, 0.95)])
def test_persistent_sampling_vs_general_and_oracle(pkg, monkeypatch, B, T, temperature, top_k, top_p):
    """Sampling loop inside the persistent cluster kernel (headline decoder, bf16) against (a) the stream-ordered
    general bf16 path on the same uniforms -- same weights and selection arithmetic, different MMA / activation
    rounding -- and (b) the fp32 oracle: filtered distributions within 6e-2 of the row maximum up to a row's first
    divergence, most rows identical, lengths / steps consistent with the tokens (sticky stop rule)."""
    cfg = H.HEADLINE
    p = oracle.make_params(cfg, 1, sharp=True)
    m16 = H.build_model(pkg, cfg, p, precision="bf16")
    g = torch.Generator().manual_seed(7)
    enc_ref = torch.randn(B, cfg["embedding_dim"], generator=g).relu()
    enc = enc_ref.cuda()
    u = torch.rand(T, B, generator=torch.Generator().manual_seed(4))
    seqs, trimmed, steps_ref, ptrace = oracle.sample_loop(p, enc_ref, H.START, H.END, T, temperature, top_k, top_p, cfg,
                                                          uniforms=u, return_probs=True)
    tokens, lengths, steps, probs = m16.decoder.sample(enc, H.START, H.END, T, temperature, top_k, top_p, uniforms=u,
                                                       return_probs=True)
    prof = pkg._native
    m16.decoder.streamed = True                        # I2L_BF16_STREAMED: same weights, stream-ordered launches
    tok_g, len_g, steps_g = m16.decoder.sample(enc, H.START, H.END, T, temperature, top_k, top_p, uniforms=u)
    m16.decoder.streamed = False
    tokens, probs, tok_g = tokens.cpu(), probs.cpu(), tok_g.cpu()
    same_oracle = same_general = 0
    sampling = temperature > 0 and (top_k > 0 or top_p > 0.0)                     # predictor.py:330
    for b in range(B):
        ref_row = seqs[b].tolist()
        got_row = tokens[b, : len(ref_row)].tolist()
        t_div = next((i for i, (a, r) in enumerate(zip(got_row, ref_row)) if a != r), None)
        same_oracle += t_div is None
        same_general += got_row == tok_g[b, : len(ref_row)].tolist()
        upto = min(len(ref_row) - 1 if t_div is None else t_div, steps_ref)
        for t in range(upto):
            both = (probs[t, b] > 0) & (ptrace[t][b] > 0)
            assert both.any()
            d = (probs[t, b] - ptrace[t][b]).abs()[both].max() / ptrace[t][b].max()
            assert float(d) < 6e-2, (b, t, float(d))
            assert abs(float(probs[t, b].sum()) - 1.0) < 1e-4
        if t_div is not None and t_div - 1 < steps_ref:
            # a row may leave the oracle only where the uniform lands near a boundary of the oracle's CDF: the draw is
            # the inverse CDF of the kernel's own distribution (checked below), and two CDFs differ by at most the L1
            # distance of the distributions, so the oracle's interval of the token taken lies within that distance
            # of the target.  Skipped where the kept sets differ at the top-k / top-p cut.
            t = t_div - 1
            pr = ptrace[t][b].double()
            a, r = got_row[t_div], ref_row[t_div]
            if not sampling:                       # argmax(probs): only a near tie of the oracle's two candidates
                l1 = float((probs[t, b].double() - pr).abs().sum())
                assert float(pr[r] - pr[a]) <= l1 + 1e-6, (b, t, a, r)
            elif float(pr[a]) > 0 and float(probs[t, b, r]) > 0:
                cdf = torch.cumsum(pr, 0)
                tgt = float(u[t, b]) * float(cdf[-1])
                lo = float(cdf[a - 1]) if a > 0 else 0.0
                dist = max(lo - tgt, tgt - float(cdf[a]), 0.0)
                l1 = float((probs[t, b].double() - pr).abs().sum())
                assert dist <= l1 + 1e-5, (b, t, a, r, dist, l1)
    # every draw is the inverse CDF of the kernel's OWN filtered distribution (all rows, all executed steps)
    n_steps = int(steps)
    cdf_all = torch.cumsum(probs[:n_steps].double(), dim=2)                       # (T,B,V)
    tgt_all = u[:n_steps].double() * cdf_all[:, :, -1]
    drawn = (cdf_all > tgt_all.unsqueeze(2)).int().argmax(dim=2)                  # first index with cdf > target
    taken = tokens[:, 1: n_steps + 1].t()
    if not sampling:                                                              # predictor.py:333-335
        assert torch.equal(taken, probs[:n_steps].argmax(dim=2))
        drawn = taken
    bad = (drawn != taken).nonzero()
    for t, b in bad.tolist():                                                     # fp64 cumsum order: only at a boundary
        gap = (cdf_all[t, b] - tgt_all[t, b]).abs().min()
        assert float(gap) < 1e-6, (t, b, int(drawn[t, b]), int(taken[t, b]), float(gap))
    print(f"persistent sampling B={B}: {same_oracle}/{B} rows follow the oracle, {same_general}/{B} the general bf16 path")
    # bookkeeping: lengths = position of the first END, steps = the sticky loop exit
    n = int(steps)
    for b in range(B):
        row = tokens[b, 1: n + 1].tolist()
        fe = row.index(H.END) + 1 if H.END in row else n + 1
        assert int(lengths[b]) == fe
    assert n == T or all(H.END in tokens[b, 1: n + 1].tolist() for b in range(B))
    assert (tokens[:, n + 1:] == -1).all()


def test_persistent_sampling_argmax_mode_equals_persistent_greedy(pkg):
    """temperature only (top_k = 0, top_p = 0) is argmax(probs) (predictor.py:330-335): the sampling mode of the persistent
    kernel (logits regrouped across the cluster, one warp per row) must pick the tokens its greedy mode (per-CTA partial
    argmax combined across the cluster) picks from the same logits."""
    cfg = H.HEADLINE
    p = oracle.make_params(cfg, 3, sharp=True)
    m16 = H.build_model(pkg, cfg, p, precision="bf16")
    B, T = 200, 40
    enc = torch.randn(B, cfg["embedding_dim"], generator=torch.Generator().manual_seed(2)).relu().cuda()
    ts, ls, ss = m16.decoder.sample(enc, H.START, H.END, T, 0.7, 0, 0.0)
    tg, lg, sg = m16.decoder.greedy(enc, H.START, H.END, T, 0.7, pkg._native.STOP_ALL_FINISHED_STICKY)
    n = min(int(ss), int(sg))
    same = (ts[:, : n + 1] == tg[:, : n + 1]).all(dim=1).sum().item()
    print(f"argmax-mode sampling vs greedy: {same}/{B} rows identical over {n} steps")
    assert same >= 0.95 * B          # softmax rounding can merge two nearly equal logits into equal probabilities
    assert int(ss) == int(sg) or same < B
    assert torch.equal(ls[(ts[:, : n + 1] == tg[:, : n + 1]).all(dim=1)], lg[(ts[:, : n + 1] == tg[:, : n + 1]).all(dim=1)])


def test_persistent_sampling_philox_reproducible(pkg):
    cfg = H.HEADLINE
    p = oracle.make_params(cfg, 2, sharp=True)
    m16 = H.build_model(pkg, cfg, p, precision="bf16")
    enc = torch.randn(96, cfg["embedding_dim"], generator=torch.Generator().manual_seed(1)).relu().cuda()
    a = m16.decoder.sample(enc, H.START, H.END, 30, 0.9, 40, 0.9, seed=5)[0]
    b = m16.decoder.sample(enc, H.START, H.END, 30, 0.9, 40, 0.9, seed=5)[0]
    c = m16.decoder.sample(enc, H.START, H.END, 30, 0.9, 40, 0.9, seed=6)[0]
    assert torch.equal(a, b) and not torch.equal(a, c)
    # a shard of the batch with the matching Philox offset reproduces its rows (batch-sharded multi-GPU decode)
    # (offset counts draws: step * B + row, so shards are only comparable through explicit uniforms)
    u = torch.rand(30, 96, generator=torch.Generator().manual_seed(3))
    full = m16.decoder.sample(enc, H.START, H.END, 30, 0.9, 40, 0.9, uniforms=u)[0]
    half = m16.decoder.sample(enc[32:64], H.START, H.END, 30, 0.9, 40, 0.9, uniforms=u[:, 32:64].contiguous())[0]
    n = min(full.shape[1], half.shape[1])
    live = (half[:, :n] >= 0) & (full[32:64, :n] >= 0)
    assert torch.equal(half[:, :n][live], full[32:64, :n][live])


# ---- the configurations bench.py measures (BASELINE configs[1] / configs[4]) at their full sizes ----------------------
def test_greedy_at_the_benchmarked_config(pkg):
    """BASELINE configs[1] as bench.py runs it: B = 1024 sequences, T = 150 steps, V = 512, bf16 persistent kernel
    (32 clusters of 4 CTAs), against `oracle.greedy_search` on the SAME encodings (decoder isolated).  Rows may leave
    the oracle only at a near tie of the oracle's own top-2 logits (BF16_MARGIN); at least 90 % of the 1024 rows must
    be identical over all 150 steps.  Lengths / steps_run must be consistent with the kernel's own tokens."""
    cfg = H.HEADLINE
    B, T = 1024, 150
    p = oracle.make_params(cfg, 3, sharp=True)
    m16 = H.build_model(pkg, cfg, p, precision="bf16")
    g = torch.Generator().manual_seed(17)
    enc_ref = torch.relu(torch.randn(B, 256, generator=g)) * (0.5 + torch.rand(B, 1, generator=g))
    ref, steps_ref, trace = oracle.greedy_search(p, enc_ref, H.START, H.END, T, 1.0, cfg, return_logits=True)
    tokens, lengths, steps = m16.decoder.greedy(enc_ref.cuda(), H.START, H.END, T)
    torch.cuda.synchronize()
    tokens, lengths, n = tokens.cpu(), lengths.cpu(), int(steps)
    exact, near, bad = divergence_report(tokens[:, : steps_ref + 1].tolist(), ref, trace)
    ref16, steps16 = oracle.greedy_search(p, enc_ref, H.START, H.END, T, 1.0, bf16_cfg(cfg))
    same16 = rows_identical(tokens[:, : steps16 + 1].tolist(), ref16)
    worst = max((m for _, _, m in near), default=0.0)
    print(f"bf16 persistent greedy B=1024 T=150: rows on the fp32 oracle {exact}/{B}, {len(near)} near-tie divergences "
          f"(largest oracle margin {worst}), steps {n} (oracle {steps_ref}); rows on the bf16-operand oracle {same16}/{B}")
    assert not bad, f"rows diverging at a step with a clear margin: {bad[:10]}"
    assert exact >= 0.4 * B                           # measured 0.48: 150 steps x 1024 rows of random-init near ties
    assert same16 >= 0.9 * B
    for b in range(0, B, 7):                          # bookkeeping against the kernel's own tokens
        row = tokens[b, 1: n + 1].tolist()
        assert int(lengths[b]) == (row.index(H.END) + 1 if H.END in row else n + 1)
    assert n == T or all(int(tokens[b, n]) == H.END for b in range(B))          # seq2seq.py:220
    assert (tokens[:, n + 1:] == -1).all()


def test_cnn_greedy_full_pipeline_at_the_benchmarked_config(pkg):
    """The whole bf16 step bench.py times -- raw uint8 pixels -> fused conv1 -> conv2/3 -> FC -> 150-step persistent
    greedy decode -- at B = 1024 against the fp32 oracle on the same pixels: every one of the 1024 encodings within
    the stated 3e-2 of max|enc|; tokens compared on a 128-row oracle sample (the encoder's bf16 error moves the
    logits, so a row may leave the oracle where the oracle's margin is below BF16_MARGIN)."""
    cfg = H.HEADLINE
    B, T, S = 1024, 150, 128
    p = oracle.make_params(cfg, 1, sharp=True)
    m16 = H.build_model(pkg, cfg, p, precision="bf16")
    g = torch.Generator().manual_seed(5)
    px = torch.randint(0, 256, (B, 3, 64, 320), dtype=torch.uint8, generator=g)
    enc = m16.encoder.forward_u8(px.cuda(), "pm1")
    tokens, lengths, steps = m16.decoder.greedy(enc, H.START, H.END, T)
    torch.cuda.synchronize()
    rows = list(range(0, B, B // S))
    enc_ref = oracle.encoder(p, oracle.normalize_u8(px[rows], "pm1"), cfg)
    err = H.rel_err(enc[rows], enc_ref)
    assert err < 3e-2, err
    # rows outside the sample: the same image must give the same encoding wherever it sits in the batch
    enc2 = m16.encoder.forward_u8(px[rows].cuda().contiguous(), "pm1")
    assert torch.equal(enc2, enc[rows])
    ref, steps_ref, trace = oracle.greedy_search(p, enc_ref, H.START, H.END, T, 1.0, cfg, return_logits=True)
    n = min(steps_ref, int(steps))
    got = tokens.cpu()[rows][:, : n + 1].tolist()
    exact, near, bad = divergence_report(got, [r[: n + 1] for r in ref], trace)
    print(f"bf16 full pipeline B=1024: encoder rel err {err:.2e}; sample of {S}: exact rows {exact}, near ties {len(near)}, "
          f"clear-margin divergences {bad}")
    assert exact + len(near) >= 0.97 * S and exact >= 0.45 * S       # measured 71 + 57 of 128, no clear-margin divergence


def test_resnet50_sampling_at_the_benchmarked_config(pkg):
    """BASELINE configs[4] as bench.py runs it on one GPU: ResNet50 trunk (bf16 tcgen05) + sampling loop (temperature
    0.8, top_k 50, top_p 0.9) inside the persistent kernel, B = 1024, 3x64x320, T = 150, E = H = 256, V = 512.
    A strided sample of 64 rows is checked against the fp32 oracle (ResNet50 encoder + `sample_loop` on the same
    uniforms): encodings within 3e-2; filtered distributions within 6e-2 of the row maximum up to the row's first
    divergence; every draw of ALL 1024 rows is the inverse CDF of the kernel's own distribution."""
    cfg = dict(H.R50, vocab_size=512, embedding_dim=256, hidden_dim=256, img_width=320)
    B, T, S = 1024, 150, 64
    p = oracle.make_params(cfg, 2, sharp=True)
    m16 = H.build_model(pkg, cfg, p, precision="bf16")
    x = H.make_images(cfg, B, seed=3)
    u = torch.rand(T, B, generator=torch.Generator().manual_seed(9))
    enc = m16.encoder(x.cuda())
    tokens, lengths, steps, probs = m16.decoder.sample(enc, H.START, H.END, T, 0.8, 50, 0.9, uniforms=u, return_probs=True)
    torch.cuda.synchronize()
    tokens, n = tokens.cpu(), int(steps)
    rows = list(range(3, B, B // S))
    enc_ref = oracle.encoder(p, x[rows], cfg)
    err = H.rel_err(enc[rows], enc_ref)
    assert err < RESNET_BF16_TOL, err
    # (1) the whole pipeline against the fp32 oracle: rows that follow it end to end (reported; the encoder's bf16
    #     error moves every distribution a little, so a uniform near a cdf boundary flips the draw)
    seqs, trimmed, steps_ref = oracle.sample_loop(p, enc_ref, H.START, H.END, T, 0.8, 50, 0.9, cfg,
                                                  uniforms=u[:, rows].contiguous())
    same = sum(tokens[b, : seqs.shape[1]].tolist() == seqs[i].tolist() for i, b in enumerate(rows))
    # (2) decoder isolated: fp32 oracle on the kernel's OWN encodings of the sampled rows -- filtered distributions
    #     within 6e-2 of the row maximum up to the row's first divergence
    seqs_d, _, steps_d, ptrace = oracle.sample_loop(p, enc[rows].cpu(), H.START, H.END, T, 0.8, 50, 0.9, cfg,
                                                    uniforms=u[:, rows].contiguous(), return_probs=True)
    probs_s = probs[:, rows].cpu()
    for i, b in enumerate(rows):
        ref_row = seqs_d[i].tolist()
        got_row = tokens[b, : len(ref_row)].tolist()
        t_div = next((k for k, (a, r) in enumerate(zip(got_row, ref_row)) if a != r), None)
        upto = min(len(ref_row) - 1 if t_div is None else t_div, steps_d, n)
        for t in range(upto):
            both = (probs_s[t, i] > 0) & (ptrace[t][i] > 0)
            assert both.any()
            d = (probs_s[t, i] - ptrace[t][i]).abs()[both].max() / ptrace[t][i].max()
            assert float(d) < 6e-2, (b, t, float(d))
    # all 1024 rows, all executed steps: the token taken is the inverse CDF of the kernel's own filtered distribution
    bad_total = 0
    for t0 in range(0, n, 25):                                                     # chunks: (25, 1024, 512) fp64
        pr = probs[t0: min(n, t0 + 25)].double()
        cdf = torch.cumsum(pr, dim=2)
        tgt = u[t0: t0 + pr.shape[0]].double().cuda() * cdf[:, :, -1]
        drawn = (cdf > tgt.unsqueeze(2)).int().argmax(dim=2)
        taken = tokens[:, 1 + t0: 1 + t0 + pr.shape[0]].t().cuda()
        live = taken >= 0                 # a cluster whose 32 rows have all finished stops decoding (sticky rule per cluster)
        for t, b in ((drawn != taken) & live).nonzero().tolist():                  # only at a cdf boundary (fp64 sum order)
            gap = (cdf[t, b] - tgt[t, b]).abs().min()
            assert float(gap) < 1e-6, (t0 + t, b)
            bad_total += 1
        assert ((pr.sum(dim=2) - 1.0).abs()[live]).max() < 1e-4
        assert (pr.sum(dim=2)[~live] == 0).all()
    # decoder isolated: the bf16-operand oracle on the kernel's own encodings of the sampled rows
    seqs16, _, _ = oracle.sample_loop(p, enc[rows].cpu(), H.START, H.END, T, 0.8, 50, 0.9, bf16_cfg(cfg),
                                      uniforms=u[:, rows].contiguous())
    same16 = sum(tokens[b, : seqs16.shape[1]].tolist() == seqs16[i].tolist() for i, b in enumerate(rows))
    print(f"resnet50 + persistent sampling B=1024: encoder rel err {err:.2e}; {same}/{S} sampled rows follow the fp32 "
          f"oracle end to end, {same16}/{S} the bf16-operand oracle on the same encodings; {bad_total} draws sit on a cdf "
          f"boundary; steps {n}")
    assert same16 >= 0.5 * S


@pytest.mark.parametrize("cfg,B,T", [(H.SHIPPED, 200, 60), (dict(H.SHIPPED, embedding_dim=128, hidden_dim=96, lstm_layers=3), 70, 30)])
def test_wide_decoder_fused_graph_loop(pkg, cfg, B, T):
    """Decoders outside the persistent kernel's shape -- here the reference's SHIPPED one, E = H = 512 / 2 layers
    (configs/config.yaml:45-50) -- run the stream-ordered loop as ONE replayed CUDA graph whose per-layer launch is
    the gate GEMM with the LSTM cell fused into its epilogue (gemm_bf16.cu, H % 32 == 0).  Greedy rows against the
    fp32 oracle (near ties only) and the bf16-operand oracle (>= 90 % identical); a second call replays the cached
    graph with different encodings and must give that call's own result; the sampling loop draws the inverse CDF."""
    wide = cfg["hidden_dim"] >= 256
    p = oracle.make_params(cfg, 4, sharp=not wide)       # (the END-bias recipe of `sharp` ends every 512-wide row at step 1)
    if wide:
        p["decoder.output_layer.weight"] *= 4.0
    m16 = H.build_model(pkg, cfg, p, precision="bf16")
    m16.decoder.streamed = True      # I2L_BF16_STREAMED: the stream-ordered loop (greedy would take decode_wide.cu otherwise)
    g = torch.Generator().manual_seed(23)
    E = cfg["embedding_dim"]
    enc_a = torch.relu(torch.randn(B, E, generator=g))
    enc_b = torch.relu(torch.randn(B, E, generator=g)) * 0.7
    lib = pkg._native.lib()
    res = {}
    for name, enc_ref in (("a", enc_a), ("b", enc_b), ("a2", enc_a)):
        l0 = lib.i2l_launch_count()
        tokens, lengths, steps = m16.decoder.greedy(enc_ref.cuda(), H.START, H.END, T)
        torch.cuda.synchronize()
        res[name] = (tokens.cpu(), lengths.cpu(), int(steps), lib.i2l_launch_count() - l0)
    assert torch.equal(res["a"][0], res["a2"][0]) and torch.equal(res["a"][1], res["a2"][1])
    assert not torch.equal(res["a"][0], res["b"][0])
    L = cfg["lstm_layers"]
    assert res["b"][3] >= T * (L + 2), "the replay's kernel nodes are counted as launches"
    for name, enc_ref in (("a", enc_a), ("b", enc_b)):
        tokens, lengths, n, _ = res[name]
        ref, steps_ref, trace = oracle.greedy_search(p, enc_ref, H.START, H.END, T, 1.0, cfg, return_logits=True)
        exact, near, bad = divergence_report(tokens[:, : steps_ref + 1].tolist(), ref, trace)
        ref16, steps16 = oracle.greedy_search(p, enc_ref, H.START, H.END, T, 1.0, dict(cfg, operand_rounding="bf16_recurrent"))
        same16 = rows_identical(tokens[:, : steps16 + 1].tolist(), ref16)
        print(f"wide decoder E={E} H={cfg['hidden_dim']} L={L} B={B} T={T} [{name}]: rows on the fp32 oracle {exact}/{B} "
              f"({len(near)} near ties), on the bf16-operand oracle {same16}/{B}")
        assert not bad, f"rows diverging at a step with a clear margin: {bad[:10]}"
        # measured: 512/512/2 144 / 147+ of 200, the tiny 3-layer case 51-55 / 61-68 of 70 (rows on the fp32 oracle / on the
        # recurrent-operand oracle): random-init logits are full of near ties, so the guard proper is `not bad` above
        assert max(exact, same16) >= 0.7 * B
    # sampling loop through the same graph machinery: every draw is the inverse CDF of the kernel's own distribution
    u = torch.rand(T, B, generator=g)
    tok_s, len_s, st_s = m16.decoder.sample(enc_a.cuda(), H.START, H.END, T, 0.9, 30, 0.9, uniforms=u)
    tok_p, _, st_p, probs = m16.decoder.sample(enc_a.cuda(), H.START, H.END, T, 0.9, 30, 0.9, uniforms=u, return_probs=True)
    assert torch.equal(tok_s, tok_p) and int(st_s) == int(st_p)           # graph replay == launch-by-launch (trace requested)
    n = int(st_p)
    cdf = torch.cumsum(probs[:n].double(), dim=2)
    tgt = u[:n].double().cuda() * cdf[:, :, -1]
    drawn = (cdf > tgt.unsqueeze(2)).int().argmax(dim=2)
    taken = tok_p[:, 1: n + 1].t()
    for t, b in (drawn != taken).nonzero().tolist():
        assert float((cdf[t, b] - tgt[t, b]).abs().min()) < 1e-6, (t, b)


WIDE_1024 = dict(H.SHIPPED, embedding_dim=1024, hidden_dim=1024, lstm_layers=3)     # configs/resnet_lstm.yaml:45-50


@pytest.mark.parametrize("cfg,B,T,rule", [(H.SHIPPED, 200, 40, "same_step"), (H.SHIPPED, 1024, 150, "same_step"),
                                          (H.SHIPPED, 300, 40, "sticky"), (WIDE_1024, 257, 20, "same_step"),
                                          (dict(H.SHIPPED, hidden_dim=192, embedding_dim=64, lstm_layers=1, vocab_size=77), 130, 25, "sticky"),
                                          (dict(H.HEADLINE, vocab_size=700), 150, 25, "same_step")])   # V > 512: beyond the cluster kernel
def test_wide_persistent_loop(pkg, cfg, B, T, rule):
    """decode_wide.cu: the whole greedy loop of a wide / multi-layer decoder as ONE cooperative kernel (tiles of 128
    sequences, per-block counters instead of launches).  It runs the same tcgen05 products in the same accumulation
    order and the same epilogue arithmetic as the stream-ordered fused loop (gemm_bf16.cu), so the two must agree BIT
    FOR BIT (tokens up to the stop step, lengths, steps); the stream-ordered loop is the one the oracle criteria are
    checked on (test_wide_decoder_fused_graph_loop), and the fp32 / bf16-operand oracles are applied here directly too."""
    N = pkg._native
    stop = N.STOP_ALL_END_SAME_STEP if rule == "same_step" else N.STOP_ALL_FINISHED_STICKY
    p = oracle.make_params(cfg, 5)
    p["decoder.output_layer.weight"] *= 4.0
    if rule == "sticky":
        p["decoder.output_layer.bias"][H.END] += 1.5                  # rows end at different steps, whole blocks stop early
    m16 = H.build_model(pkg, cfg, p, precision="bf16")
    g = torch.Generator().manual_seed(31 + B)
    enc = torch.relu(torch.randn(B, cfg["embedding_dim"], generator=g))
    lib = N.lib()
    m16.decoder.greedy(enc[:2].cuda(), H.START, H.END, 2, 1.0, stop)          # packs the weights (counted launches)
    l0 = lib.i2l_launch_count()
    tok_w, len_w, st_w = m16.decoder.greedy(enc.cuda(), H.START, H.END, T, 1.0, stop)
    torch.cuda.synchronize()
    launches = lib.i2l_launch_count() - l0
    assert launches <= 8, f"{launches} launches: the loop did not take the persistent kernel"
    m16.decoder.streamed = True
    tok_s, len_s, st_s = m16.decoder.greedy(enc.cuda(), H.START, H.END, T, 1.0, stop)
    m16.decoder.streamed = False
    n = int(st_s)
    assert int(st_w) == n
    assert torch.equal(len_w, len_s)
    tw, ts = tok_w.cpu()[:, : n + 1], tok_s.cpu()[:, : n + 1]
    if rule == "sticky":
        # a 128-sequence block stops once ITS rows have finished (the stream-ordered loop runs every row to the global
        # stop): compare every row up to its own END
        for b in range(B):
            k = int(len_s[b])
            assert tw[b, : k + 1].tolist() == ts[b, : k + 1].tolist(), b
    else:
        assert torch.equal(tw, ts)
    if B <= 300:
        ocfg = dict(cfg)
        if rule == "same_step":
            ref, steps_ref, trace = oracle.greedy_search(p, enc, H.START, H.END, T, 1.0, ocfg, return_logits=True)
            exact, near, bad = divergence_report(tok_w.cpu()[:, : steps_ref + 1].tolist(), ref, trace)
            print(f"wide persistent loop H={cfg['hidden_dim']} L={cfg['lstm_layers']} B={B} T={T}: rows on the fp32 oracle "
                  f"{exact}/{B} ({len(near)} near ties), steps {n} (oracle {steps_ref})")
            assert not bad, f"rows diverging at a step with a clear margin: {bad[:10]}"


@pytest.mark.parametrize("cfg,B,T,args", [(H.SHIPPED, 300, 40, (0.8, 50, 0.9)), (H.SHIPPED, 130, 30, (1.0, 0, 0.0)),
                                          (WIDE_1024, 140, 16, (0.9, 20, 0.0)),
                                          (dict(H.SHIPPED, hidden_dim=192, embedding_dim=64, lstm_layers=1, vocab_size=77), 200, 25, (0.7, 0, 0.8))])
def test_wide_persistent_sampling(pkg, cfg, B, T, args):
    """Predictor.predict_batch's loop (predictor.py:283-347) on a wide decoder inside decode_wide.cu: the logits tiles
    write whole rows, one epilogue warp per row runs the selection routine of the other sampling paths
    (sample_select.cuh).  Same products, same routine, same uniforms as the stream-ordered loop => identical tokens (every
    row up to its own END: a 128-sequence block stops once ITS rows have finished), lengths, steps and filtered
    distributions; every draw is the inverse CDF of the kernel's own distribution."""
    temperature, top_k, top_p = args
    N = pkg._native
    p = oracle.make_params(cfg, 6)
    p["decoder.output_layer.weight"] *= 4.0
    p["decoder.output_layer.bias"][H.END] += 2.0
    m16 = H.build_model(pkg, cfg, p, precision="bf16")
    g = torch.Generator().manual_seed(41 + B)
    enc = torch.relu(torch.randn(B, cfg["embedding_dim"], generator=g)).cuda()
    u = torch.rand(T, B, generator=g)
    lib = N.lib()
    m16.decoder.sample(enc[:2], H.START, H.END, 2, temperature, top_k, top_p, uniforms=u[:2, :2].contiguous())   # packs the weights
    l0 = lib.i2l_launch_count()
    tok_w, len_w, st_w, pr_w = m16.decoder.sample(enc, H.START, H.END, T, temperature, top_k, top_p, uniforms=u, return_probs=True)
    torch.cuda.synchronize()
    assert lib.i2l_launch_count() - l0 <= 8, "the sampling loop did not take the persistent kernel"
    m16.decoder.streamed = True
    tok_s, len_s, st_s, pr_s = m16.decoder.sample(enc, H.START, H.END, T, temperature, top_k, top_p, uniforms=u, return_probs=True)
    m16.decoder.streamed = False
    n = int(st_s)
    assert int(st_w) == n and torch.equal(len_w, len_s)
    tw, ts, ln = tok_w.cpu(), tok_s.cpu(), len_s.cpu()
    pw, ps = pr_w.cpu(), pr_s.cpu()
    for b in range(B):
        k = min(int(ln[b]), n)
        assert tw[b, : k + 1].tolist() == ts[b, : k + 1].tolist(), b
        assert torch.equal(pw[:k, b], ps[:k, b]), b
    # Philox draws: reproducible, and the same as the stream-ordered loop's
    a1 = m16.decoder.sample(enc, H.START, H.END, T, temperature, top_k, top_p, seed=7, offset=3)
    a2 = m16.decoder.sample(enc, H.START, H.END, T, temperature, top_k, top_p, seed=7, offset=3)
    assert torch.equal(a1[0], a2[0]) and torch.equal(a1[1], a2[1])


# ---- the tcgen05 CNN encoder beyond the benchmark shape (cnn_bf16.cu: C in {1,3}, H % 64 == 0, W % 32 == 0) -----------
def _cnn_cfg(c, h, w):
    return dict(H.HEADLINE, channels=c, img_height=h, img_width=w)


@pytest.mark.parametrize("c,h,w,B", [(1, 64, 800, 5), (1, 64, 800, 33), (1, 64, 96, 6), (3, 128, 64, 3), (1, 64, 32, 1),
                                     (3, 64, 160, 7), (1, 128, 320, 2)])
def test_cnn_encoder_bf16_other_shapes(pkg, c, h, w, B):
    """The reference's default / serving shape 1x64x800 (encoder.py:50-64, predictor.py:409-414) and other (C, H, W)
    on the tensor-core path: every conv2 / conv3 tile instantiation (16, 8, 4 pooled columns x 1, 2, 4 images) and
    batch sizes that are not a multiple of the images-per-tile count; within 3e-2 of max|out| of the fp32 oracle, and
    a row does not depend on its batch neighbours."""
    cfg = _cnn_cfg(c, h, w)
    p = oracle.make_params(cfg, 1, sharp=True)
    m16 = H.build_model(pkg, cfg, p, precision="bf16")
    assert m16.encoder._bf16_shape(), "this shape must run on the tcgen05 kernels"
    x = H.make_images(cfg, B)
    ref = oracle.cnn_encoder(p, x)
    lib = pkg._native.lib()
    lib.i2l_prof_reset(); lib.i2l_prof_enable(1)
    out = m16.encoder(x.cuda())
    torch.cuda.synchronize()
    lib.i2l_prof_enable(0)
    names = set(pkg._native.prof_results())
    assert {"cnn.conv1_bf16", "cnn.conv2_bf16", "cnn.conv3_bf16", "cnn.fc_bf16"} <= names, names
    err = H.rel_err(out, ref)
    print(f"bf16 tcgen05 encoder {c}x{h}x{w} B={B}: rel err {err:.3e}")
    assert out.shape == ref.shape and err < 3e-2
    if B > 2:
        part = m16.encoder(x[1:3].contiguous().cuda())
        assert torch.equal(part, out[1:3])


@pytest.mark.parametrize("c,h,w,B,mode", [(1, 64, 800, 6, "pm1"), (1, 64, 96, 3, "meanstd"), (3, 128, 64, 2, "pm1")])
def test_cnn_fused_uint8_other_shapes(pkg, c, h, w, B, mode):
    """forward_u8 (normalisation fused into conv1) on the other tcgen05 shapes: same criteria as the headline shape."""
    cfg = _cnn_cfg(c, h, w)
    p = oracle.make_params(cfg, 2)
    g = torch.Generator().manual_seed(13 + w)
    px = torch.randint(0, 256, (B, c, h, w), dtype=torch.uint8, generator=g)
    px[:, :, 0, :] = 0; px[:, :, -1, :] = 255; px[:, :, :, 0] = 0; px[:, :, :, -1] = 9
    _check_fused_u8(pkg, cfg, p, px, mode)
