"""CPU restatement (fp32, torch CPU functional ops) of the reference inference path.

TEST INFRASTRUCTURE -- see ``oracle/__init__.py``.  Every function cites the
reference lines it restates (paths relative to the reference checkout).  The
arithmetic of the reference lives in torch / torchvision (un-vendored deps,
``pyproject.toml:14-15``; de-facto pin = torch 2.11.0 / torchvision 0.26.0 of
this image), so the restatement calls the same ATen CPU kernels
(``F.conv2d``, ``F.linear`` ...) and writes out by hand what the reference
delegates to ``nn.LSTM`` / ``torchvision.models.resnet*``.

Parameters travel as a flat ``dict[str, Tensor]`` whose keys are exactly the
reference ``Seq2SeqModel.state_dict()`` keys.
"""
from __future__ import annotations

import math
from typing import Dict, List, Optional, Sequence, Tuple

import torch
import torch.nn.functional as F

Params = Dict[str, torch.Tensor]

__all__ = [
    "make_params", "cnn_encoder", "resnet_encoder", "encoder", "attention",
    "lstm_step", "decode_step", "greedy_search", "inference_postprocess",
    "filter_probs", "sample_loop", "beam_search", "beam_search_batched",
    "RESNET_LAYERS", "trim_at_end", "inverse_cdf_draw", "normalize_u8", "decoder_forward", "seq2seq_forward",
]

RESNET_LAYERS = {
    "resnet18": ("basic", (2, 2, 2, 2)),
    "resnet34": ("basic", (3, 4, 6, 3)),
    "resnet50": ("bottleneck", (3, 4, 6, 3)),
    "resnet101": ("bottleneck", (3, 4, 23, 3)),
    "resnet152": ("bottleneck", (3, 8, 36, 3)),
}


IMAGENET_MEAN = (0.485, 0.456, 0.406)
IMAGENET_STD = (0.229, 0.224, 0.225)


def normalize_u8(pixels: torch.Tensor, mode: str = "pm1") -> torch.Tensor:
    """Pixel arithmetic of `load_image` (data/utils.py:53-80) and the PIL branch of
    `Predictor._prepare_image` (training/predictor.py:441-446): uint8 (.., C, H, W) ->
    float32 / 255.0, then `* 2.0 - 1.0` ("pm1": grayscale utils.py:72-74, PIL predictor.py:446)
    or ImageNet `(x - mean) / std` ("meanstd": RGB, utils.py:76-79)."""
    t = pixels.float() / 255.0
    if mode == "pm1":
        return t * 2.0 - 1.0
    C = pixels.shape[-3]
    mean = torch.tensor(IMAGENET_MEAN[:C]).view(-1, 1, 1)
    std = torch.tensor(IMAGENET_STD[:C]).view(-1, 1, 1)
    return (t - mean) / std


# --------------------------------------------------------------------------
# deterministic synthetic parameters (test / bench inputs, not reference code)
# --------------------------------------------------------------------------
def _uniform(gen, shape, bound):
    return (torch.rand(shape, generator=gen, dtype=torch.float32) * 2.0 - 1.0) * bound


def make_params(cfg: dict, seed: int = 0, sharp: bool = False) -> Params:
    """Seeded random-init parameters with the reference's state_dict keys and
    PyTorch-default-like scales (U(-1/sqrt(fan_in), 1/sqrt(fan_in))).

    ``sharp=True`` rescales the output layer / embedding so that the greedy
    sequences depend on the image and END appears at varied steps (SURVEY F13:
    the default init is nearly input independent and never emits END).

    cfg keys: model_type ('cnn_lstm'|'resnet_lstm'), vocab_size, embedding_dim,
    hidden_dim, lstm_layers, attention, img_height, img_width, channels,
    conv_filters, kernel_size, model_name.
    """
    g = torch.Generator().manual_seed(seed)
    p: Params = {}
    E = cfg.get("embedding_dim", 256)
    H = cfg.get("hidden_dim", 256)
    L = cfg.get("lstm_layers", 1)
    V = cfg["vocab_size"]
    mt = cfg.get("model_type", "cnn_lstm")
    if mt == "cnn_lstm":
        cin = cfg.get("channels", 3)
        ks = cfg.get("kernel_size", 3)
        filters = cfg.get("conv_filters", [32, 64, 128])
        h, w = cfg["img_height"], cfg["img_width"]
        for i, f in enumerate(filters):
            bound = 1.0 / math.sqrt(cin * ks * ks)
            p[f"encoder.cnn_layers.{3 * i}.weight"] = _uniform(g, (f, cin, ks, ks), bound)
            p[f"encoder.cnn_layers.{3 * i}.bias"] = _uniform(g, (f,), bound)
            cin = f
            h, w = h // 2, w // 2
        flat = cin * h * w
        bound = 1.0 / math.sqrt(flat)
        p["encoder.embedding_layer.weight"] = _uniform(g, (E, flat), bound)
        p["encoder.embedding_layer.bias"] = _uniform(g, (E,), bound)
    elif mt == "resnet_lstm":
        kind, layers = RESNET_LAYERS[cfg.get("model_name", "resnet50")]

        def conv(name, co, ci, k):
            bound = math.sqrt(3.0) * math.sqrt(2.0 / (ci * k * k))  # variance of kaiming_normal
            p[name + ".weight"] = _uniform(g, (co, ci, k, k), bound)

        def bn(name, c):
            p[name + ".weight"] = 0.5 + torch.rand(c, generator=g) * 0.5
            p[name + ".bias"] = _uniform(g, (c,), 0.1)
            p[name + ".running_mean"] = _uniform(g, (c,), 0.1)
            p[name + ".running_var"] = 0.5 + torch.rand(c, generator=g)
            p[name + ".num_batches_tracked"] = torch.tensor(0, dtype=torch.long)

        conv("encoder.resnet.0", 64, 3, 7)
        bn("encoder.resnet.1", 64)
        inpl = 64
        exp = 1 if kind == "basic" else 4
        for li, nblk in enumerate(layers):
            planes = 64 * (2 ** li)
            for b in range(nblk):
                stride = 2 if (li > 0 and b == 0) else 1
                pre = f"encoder.resnet.{4 + li}.{b}"
                if kind == "basic":
                    conv(pre + ".conv1", planes, inpl, 3); bn(pre + ".bn1", planes)
                    conv(pre + ".conv2", planes, planes, 3); bn(pre + ".bn2", planes)
                else:
                    conv(pre + ".conv1", planes, inpl, 1); bn(pre + ".bn1", planes)
                    conv(pre + ".conv2", planes, planes, 3); bn(pre + ".bn2", planes)
                    conv(pre + ".conv3", planes * 4, planes, 1); bn(pre + ".bn3", planes * 4)
                if stride != 1 or inpl != planes * exp:
                    conv(pre + ".downsample.0", planes * exp, inpl, 1)
                    bn(pre + ".downsample.1", planes * exp)
                inpl = planes * exp
        bound = 1.0 / math.sqrt(inpl)
        p["encoder.embedding_layer.weight"] = _uniform(g, (E, inpl), bound)
        p["encoder.embedding_layer.bias"] = _uniform(g, (E,), bound)
    else:
        raise ValueError(f"Invalid model type: {mt}")

    p["decoder.embedding.weight"] = torch.randn((V, E), generator=g)
    for l in range(L):
        insz = 2 * E if l == 0 else H
        bound = 1.0 / math.sqrt(H)
        p[f"decoder.lstm.weight_ih_l{l}"] = _uniform(g, (4 * H, insz), bound)
        p[f"decoder.lstm.weight_hh_l{l}"] = _uniform(g, (4 * H, H), bound)
        p[f"decoder.lstm.bias_ih_l{l}"] = _uniform(g, (4 * H,), bound)
        p[f"decoder.lstm.bias_hh_l{l}"] = _uniform(g, (4 * H,), bound)
    if cfg.get("attention", True):
        bound = 1.0 / math.sqrt(H + E)
        p["decoder.attention.attn.weight"] = _uniform(g, (H, H + E), bound)
        p["decoder.attention.attn.bias"] = _uniform(g, (H,), bound)
        p["decoder.attention.v.weight"] = _uniform(g, (1, H), 1.0 / math.sqrt(H))
    bound = 1.0 / math.sqrt(H)
    p["decoder.output_layer.weight"] = _uniform(g, (V, H), bound)
    p["decoder.output_layer.bias"] = _uniform(g, (V,), bound)
    if sharp:
        # empirically tuned so greedy sequences differ per image and END (id 2) lands at varied steps
        small = H < 100
        p["encoder.embedding_layer.weight"] *= 4.0
        p["decoder.output_layer.weight"] *= 8.0 if small else 4.0
        for l in range(L):
            p[f"decoder.lstm.weight_ih_l{l}"] *= 4.0
            p[f"decoder.lstm.weight_hh_l{l}"] *= 3.0 if small else 1.0
        p["decoder.output_layer.bias"][2] += 1.5
    return p


# --------------------------------------------------------------------------
# encoders
# --------------------------------------------------------------------------
def cnn_encoder(p: Params, x: torch.Tensor, pool_size: int = 2) -> torch.Tensor:
    """`CNNEncoder.forward` (img2latex/model/encoder.py:111-129; layer stack
    74-95): N x [Conv2d 'same' -> ReLU -> MaxPool2d(pool)] -> Flatten (NCHW
    order) -> Linear -> ReLU."""
    i = 0
    while f"encoder.cnn_layers.{3 * i}.weight" in p:
        w = p[f"encoder.cnn_layers.{3 * i}.weight"]
        b = p[f"encoder.cnn_layers.{3 * i}.bias"]
        x = F.conv2d(x, w, b, padding=w.shape[-1] // 2)       # encoder.py:80-87
        x = F.relu(x)                                          # encoder.py:88
        x = F.max_pool2d(x, kernel_size=pool_size)             # encoder.py:91
        i += 1
    x = x.flatten(1)                                           # encoder.py:125
    x = F.linear(x, p["encoder.embedding_layer.weight"], p["encoder.embedding_layer.bias"])
    return F.relu(x)                                           # encoder.py:126-127


def _bn(p: Params, name: str, x: torch.Tensor) -> torch.Tensor:
    # eval-mode BatchNorm2d (torchvision default eps=1e-5); predictor.py:55 sets .eval()
    return F.batch_norm(x, p[name + ".running_mean"], p[name + ".running_var"],
                        p[name + ".weight"], p[name + ".bias"], training=False, eps=1e-5)


def resnet_encoder(p: Params, x: torch.Tensor, model_name: str = "resnet50") -> torch.Tensor:
    """`ResNetEncoder.forward` (encoder.py:231-249): torchvision ResNet trunk
    without `fc` (encoder.py:198-199) -> Flatten -> Linear -> ReLU.  The trunk
    (torchvision/models/resnet.py, v0.26: BasicBlock / Bottleneck with the
    stride on the 3x3 conv) is written out by hand."""
    if model_name not in RESNET_LAYERS:
        raise ValueError(f"Invalid ResNet model name: {model_name}")   # encoder.py:195-196
    kind, layers = RESNET_LAYERS[model_name]
    pre = "encoder.resnet"
    x = F.conv2d(x, p[f"{pre}.0.weight"], None, stride=2, padding=3)
    x = F.relu(_bn(p, f"{pre}.1", x))
    x = F.max_pool2d(x, kernel_size=3, stride=2, padding=1)
    for li, nblk in enumerate(layers):
        for b in range(nblk):
            stride = 2 if (li > 0 and b == 0) else 1
            q = f"{pre}.{4 + li}.{b}"
            idt = x
            if kind == "basic":
                o = F.conv2d(x, p[q + ".conv1.weight"], None, stride=stride, padding=1)
                o = F.relu(_bn(p, q + ".bn1", o))
                o = F.conv2d(o, p[q + ".conv2.weight"], None, stride=1, padding=1)
                o = _bn(p, q + ".bn2", o)
            else:
                o = F.conv2d(x, p[q + ".conv1.weight"], None)
                o = F.relu(_bn(p, q + ".bn1", o))
                o = F.conv2d(o, p[q + ".conv2.weight"], None, stride=stride, padding=1)
                o = F.relu(_bn(p, q + ".bn2", o))
                o = F.conv2d(o, p[q + ".conv3.weight"], None)
                o = _bn(p, q + ".bn3", o)
            if (q + ".downsample.0.weight") in p:
                idt = F.conv2d(x, p[q + ".downsample.0.weight"], None, stride=stride)
                idt = _bn(p, q + ".downsample.1", idt)
            x = F.relu(o + idt)
    x = F.adaptive_avg_pool2d(x, 1).flatten(1)                 # encoder.py:242-245
    x = F.linear(x, p["encoder.embedding_layer.weight"], p["encoder.embedding_layer.bias"])
    return F.relu(x)                                           # encoder.py:246-247


def encoder(p: Params, x: torch.Tensor, cfg: dict) -> torch.Tensor:
    """Encoder dispatch of `Seq2SeqModel.__init__` (seq2seq.py:57-80)."""
    mt = cfg.get("model_type", "cnn_lstm")
    if mt == "cnn_lstm":
        return cnn_encoder(p, x, cfg.get("pool_size", 2))
    if mt == "resnet_lstm":
        return resnet_encoder(p, x, cfg.get("model_name", "resnet50"))
    raise ValueError(f"Invalid model type: {mt}. Expected 'cnn_lstm' or 'resnet_lstm'.")


# --------------------------------------------------------------------------
# decoder step
# --------------------------------------------------------------------------
def attention(attn_w, attn_b, v_w, hidden: torch.Tensor, encoder_outputs: torch.Tensor) -> torch.Tensor:
    """`Attention.forward` (decoder.py:312-343): hidden (B,1,H), encoder_outputs
    (B,L,E) -> context (B,1,E)."""
    src_len = encoder_outputs.shape[1]
    hid = hidden.repeat(1, src_len, 1)                                         # decoder.py:329
    energy = torch.tanh(F.linear(torch.cat((hid, encoder_outputs), dim=2), attn_w, attn_b))  # :332
    att = F.linear(energy, v_w).squeeze(2)                                     # :335
    wts = F.softmax(att, dim=1).unsqueeze(1)                                   # :338
    return torch.bmm(wts, encoder_outputs)                                     # :341


def _r16(t: torch.Tensor) -> torch.Tensor:
    """round to bf16 and back (round-to-nearest-even, what the kernels' operand conversion does)"""
    return t.bfloat16().float()


def lstm_step(p: Params, x: torch.Tensor, h: torch.Tensor, c: torch.Tensor, L: int, bf16_operands=False):
    """One time step of `nn.LSTM(batch_first=True)` (decoder.py:76-82, called at
    :277) written out: gates (i,f,g,o) = x W_ih^T + b_ih + h W_hh^T + b_hh;
    c' = sig(f) c + sig(i) tanh(g); h' = sig(o) tanh(c').  x (B,In), h/c (L,B,H).
    Inter-layer dropout is inactive in eval (decoder.py:81).

    ``bf16_operands`` (cfg["operand_rounding"] == "bf16", NOT the reference's arithmetic): the same recurrence with
    the operands of the tensor-core GEMMs rounded to bf16 exactly where precision="bf16" of the CUDA path rounds them
    -- h, W_hh, the context half of layer 0's input (enc, W_ih[:, E:]), deeper layers' inputs and W_ih -- fp32
    accumulation, fp32 token term (emb W_ih[:, :E]^T), fp32 biases, cell state and activations.  Lets the tests
    separate "bf16 operand rounding" (inherent to the mode) from kernel error: against THIS variant the bf16 kernels
    must agree almost everywhere, against the fp32 restatement only up to near ties.
    ``bf16_operands="recurrent"`` (cfg["operand_rounding"] == "bf16_recurrent"): the stream-ordered path's variant --
    it computes the whole layer-0 input term (token AND context half) in fp32 and rounds only the recurrent / deeper
    products."""
    hs, cs = [], []
    inp = x
    for l in range(L):
        w_ih, w_hh = p[f"decoder.lstm.weight_ih_l{l}"], p[f"decoder.lstm.weight_hh_l{l}"]
        if bf16_operands:
            if l == 0 and bf16_operands == "recurrent":
                gi = F.linear(inp, w_ih)
            elif l == 0:
                E = w_ih.shape[1] // 2
                gi = F.linear(inp[:, :E], w_ih[:, :E]) + F.linear(_r16(inp[:, E:]), _r16(w_ih[:, E:]))
            else:
                gi = F.linear(_r16(inp), _r16(w_ih))
            gates = (gi + p[f"decoder.lstm.bias_ih_l{l}"] + p[f"decoder.lstm.bias_hh_l{l}"]
                     + F.linear(_r16(h[l]), _r16(w_hh)))
            i, f, g, o = gates.chunk(4, dim=1)
            cn = torch.sigmoid(f) * c[l] + torch.sigmoid(i) * torch.tanh(g)
            hn = torch.sigmoid(o) * torch.tanh(cn)
            hs.append(hn); cs.append(cn)
            inp = hn
            continue
        gates = (F.linear(inp, w_ih, p[f"decoder.lstm.bias_ih_l{l}"])
                 + F.linear(h[l], w_hh, p[f"decoder.lstm.bias_hh_l{l}"]))
        i, f, g, o = gates.chunk(4, dim=1)
        cn = torch.sigmoid(f) * c[l] + torch.sigmoid(i) * torch.tanh(g)
        hn = torch.sigmoid(o) * torch.tanh(cn)
        hs.append(hn); cs.append(cn)
        inp = hn
    return inp, torch.stack(hs), torch.stack(cs)


def decode_step(p: Params, encoder_output: torch.Tensor, input_token: torch.Tensor,
                hidden: Optional[Tuple[torch.Tensor, torch.Tensor]], cfg: dict):
    """`LSTMDecoder.decode_step` (decoder.py:197-284).  encoder_output (B,E),
    input_token (B,1) int64, hidden None | ((L,B,H),(L,B,H)) -> logits (B,1,V),
    (h,c)."""
    L = cfg.get("lstm_layers", 1)
    H = cfg.get("hidden_dim", 256)
    B = input_token.shape[0]
    emb = F.embedding(input_token, p["decoder.embedding.weight"])              # decoder.py:214
    if hidden is None:                                                         # :231-244 / :253-266
        hidden = (torch.zeros(L, B, H), torch.zeros(L, B, H))
    h, c = hidden
    if cfg.get("attention", True):
        ctx = attention(p["decoder.attention.attn.weight"], p["decoder.attention.attn.bias"],
                        p["decoder.attention.v.weight"], h[-1].unsqueeze(1),
                        encoder_output.unsqueeze(1))                           # :271  (src_len == 1)
    else:
        ctx = encoder_output.unsqueeze(1)                                      # :218
    x = torch.cat([emb, ctx], dim=2).squeeze(1)                                # :228 / :274
    if cfg.get("operand_rounding") in ("bf16", "bf16_recurrent"):   # see lstm_step
        top, hn, cn = lstm_step(p, x, h, c, L, bf16_operands="recurrent" if cfg["operand_rounding"] == "bf16_recurrent" else True)
        logits = F.linear(_r16(top), _r16(p["decoder.output_layer.weight"]), p["decoder.output_layer.bias"])
        return logits.unsqueeze(1), (hn, cn)
    top, hn, cn = lstm_step(p, x, h, c, L)                                     # :247 / :277
    logits = F.linear(top, p["decoder.output_layer.weight"], p["decoder.output_layer.bias"])  # :250/:280
    return logits.unsqueeze(1), (hn, cn)


def decoder_forward(p: Params, encoder_output: torch.Tensor, target_sequence: torch.Tensor, cfg: dict,
                    hidden: Optional[Tuple[torch.Tensor, torch.Tensor]] = None) -> torch.Tensor:
    """`LSTMDecoder.forward` in eval mode (decoder.py:100-195: dropout is the identity): the teacher-forced pass
    over a known token sequence.  encoder_output (B,E), target_sequence (B,T) int64 -> logits (B,T,V).  Both
    branches of the reference (one nn.LSTM call over the whole sequence without attention, 121-144; a per-step
    loop with attention over the single encoder vector, 145-193) are the recurrence of `decode_step`."""
    outs = []
    for t in range(target_sequence.shape[1]):
        logits, hidden = decode_step(p, encoder_output, target_sequence[:, t:t + 1], hidden, cfg)
        outs.append(logits)
    return torch.cat(outs, dim=1)


def seq2seq_forward(p: Params, images: torch.Tensor, target_sequences: torch.Tensor, cfg: dict) -> torch.Tensor:
    """`Seq2SeqModel.forward` (seq2seq.py:98-122): encoder, then the decoder on target_sequences[:, :-1]."""
    return decoder_forward(p, encoder(p, images, cfg), target_sequences[:, :-1], cfg)


# --------------------------------------------------------------------------
# decode loops
# --------------------------------------------------------------------------
def greedy_search(p: Params, encoder_output: torch.Tensor, start: int, end: int,
                  max_length: int, temperature: float, cfg: dict,
                  return_logits: bool = False):
    """`Seq2SeqModel._greedy_search` (seq2seq.py:192-232) up to (not including)
    the B==1 post-processing: returns raw per-row lists incl. START, and the
    number of loop iterations executed.  top_k/top_p are ignored by the
    reference here; stop iff ALL rows emit END in the same step (:220)."""
    B = encoder_output.shape[0]
    tok = torch.full((B, 1), start, dtype=torch.long)
    hidden = None
    seqs = [[start] for _ in range(B)]
    trace = []
    steps = 0
    for _ in range(max_length):
        out, hidden = decode_step(p, encoder_output, tok, hidden, cfg)
        logits = out.squeeze(1)
        if temperature != 1.0:
            logits = logits / temperature                                      # :213-214
        if return_logits:
            trace.append(logits.clone())
        nxt = torch.argmax(logits, dim=-1)                                     # :215
        tok = nxt.unsqueeze(1)
        for i in range(B):
            seqs[i].append(int(nxt[i]))
        steps += 1
        if all(t == end for t in nxt.tolist()):                                # :220
            break
    if return_logits:
        return seqs, steps, trace
    return seqs, steps


def inference_postprocess(seqs: List[List[int]], start: int, end: int):
    """Tail of `_greedy_search` (seq2seq.py:223-232): B==1 strips START and cuts
    at END; B>1 returns the raw lists."""
    if len(seqs) != 1:
        return seqs
    s = seqs[0]
    if s and s[0] == start:
        s = s[1:]
    if end in s:
        s = s[: s.index(end)]
    return s


def trim_at_end(seq: Sequence[int], end: int) -> List[int]:
    """`predict_batch` trim (predictor.py:350-358): cut at first END, exclusive."""
    seq = list(seq)
    return seq[: seq.index(end)] if end in seq else seq


def filter_probs(logits: torch.Tensor, temperature: float, top_k: int, top_p: float) -> torch.Tensor:
    """predictor.py:295-327: temperature, softmax, top-k threshold mask
    (`probs < kth`, ties kept), renorm, nucleus on descending sort with the
    mask shifted right by one (index 0 always kept), renorm."""
    if temperature != 1.0:
        logits = logits / temperature                                          # :295-296
    probs = torch.softmax(logits, dim=-1)                                      # :297
    if top_k > 0:
        k = min(top_k, probs.size(-1))                                         # :300
        kth, _ = torch.topk(probs, k, dim=-1)
        kth = kth[:, -1, None]                                                 # :302-304
        probs = probs.masked_fill(probs < kth, 0.0)                            # :305-306
        s = probs.sum(dim=-1, keepdim=True)
        if torch.any(s > 0):                                                   # :308
            probs = probs / s
    if top_p > 0.0:
        # stable descending sort: ties resolved lowest-index-first (documented
        # near-tie policy; torch.sort(descending=True) gives no guarantee).
        sp, si = torch.sort(probs, descending=True, stable=True)               # :312-314
        cum = torch.cumsum(sp, dim=-1)                                         # :315
        rem = cum > top_p                                                      # :316
        rem[:, 1:] = rem[:, :-1].clone()                                       # :317-319
        rem[:, 0] = False                                                      # :320
        mask = rem.scatter(-1, si, rem)                                        # :321-323
        probs = probs.masked_fill(mask, 0.0)                                   # :324
        s = probs.sum(dim=-1, keepdim=True)
        if torch.any(s > 0):                                                   # :326
            probs = probs / s
    return probs


def inverse_cdf_draw(probs: torch.Tensor, u: torch.Tensor) -> torch.Tensor:
    """Restated draw standing in for `torch.multinomial(probs, 1)`
    (predictor.py:331), whose RNG stream a custom kernel cannot reproduce:
    token = first index j (vocab order) with cdf_j > u * cdf_last, the cdf
    accumulated in float64.  u in [0,1), shape (B,)."""
    cdf = torch.cumsum(probs.double(), dim=-1)
    tgt = (u.double() * cdf[:, -1]).unsqueeze(1)
    idx = (cdf > tgt).int().argmax(dim=-1)
    none = ~(cdf > tgt).any(dim=-1)
    if none.any():  # u*total == total can only happen through rounding: last positive prob
        lastpos = probs.shape[1] - 1 - (probs.flip(-1) > 0).int().argmax(dim=-1)
        idx = torch.where(none, lastpos, idx)
    return idx


def sample_loop(p: Params, encoder_output: torch.Tensor, start: int, end: int, max_length: int,
                temperature: float, top_k: int, top_p: float, cfg: dict,
                uniforms: Optional[torch.Tensor] = None, return_probs: bool = False):
    """Batched loop of `Predictor.predict_batch` (predictor.py:283-347) and its
    trim (350-360).  Samples iff temperature > 0 and (top_k > 0 or top_p > 0)
    (:330) using `inverse_cdf_draw` with uniforms[t, b]; else argmax(probs).
    Sticky `finished`; break iff all finished (:343-347).
    Returns (sequences (B, steps+1) int64, trimmed lists, steps[, probs trace])."""
    B = encoder_output.shape[0]
    seqs = torch.full((B, 1), start, dtype=torch.long)
    finished = torch.zeros(B, dtype=torch.bool)
    hidden = None
    steps = 0
    ptrace = []
    for t in range(max_length):
        out, hidden = decode_step(p, encoder_output, seqs[:, -1:], hidden, cfg)
        probs = filter_probs(out.squeeze(1), temperature, top_k, top_p)
        if return_probs:
            ptrace.append(probs.clone())
        if temperature > 0 and (top_k > 0 or top_p > 0.0):                     # :330
            nxt = inverse_cdf_draw(probs, uniforms[t]).unsqueeze(1)
        else:
            nxt = torch.argmax(probs, dim=-1, keepdim=True)                    # :333-335
        seqs = torch.cat([seqs, nxt], dim=1)                                   # :338-340
        finished = finished | (nxt.squeeze(1) == end)                          # :343
        steps += 1
        if torch.all(finished):                                                # :346
            break
    trimmed = [trim_at_end(r.tolist(), end) for r in seqs]                     # :350-358
    if return_probs:
        return seqs, trimmed, steps, ptrace
    return seqs, trimmed, steps


def beam_search(p: Params, encoder_output: torch.Tensor, start: int, end: int, max_length: int,
                beam_size: int, cfg: dict, return_trace: bool = False, return_cands: bool = False):
    """`Seq2SeqModel._beam_search` (seq2seq.py:234-298) for ONE image
    (encoder_output (1,E)).  Scores accumulate as Python floats (double) from
    fp32 log-probs; finished beams retire to `completed` when next visited;
    stable sort; first-wins max; no length normalisation."""
    assert encoder_output.shape[0] == 1
    beams = [{"tokens": [start], "hidden": None, "score": 0.0}]
    completed = []
    trace = []
    all_cands = []
    for _ in range(max_length):
        cands = []
        for bi, beam in enumerate(beams):
            last = beam["tokens"][-1]
            if last == end:                                                    # :258-260
                completed.append(beam)
                continue
            tok = torch.tensor([[last]], dtype=torch.long)
            out, nh = decode_step(p, encoder_output, tok, beam["hidden"], cfg)  # :262-264
            logp = torch.log_softmax(out.squeeze(1), dim=-1).squeeze(0)        # :266
            tp, ti = torch.topk(logp, beam_size)                               # :267
            for lp, ix in zip(tp.tolist(), ti.tolist()):                       # :268-275
                cands.append({"tokens": beam["tokens"] + [ix], "hidden": nh,
                              "score": beam["score"] + lp, "parent": bi})
        if not cands:                                                          # :276-277
            break
        cands = sorted(cands, key=lambda b: b["score"], reverse=True)          # :279
        all_cands.append([(b["parent"], b["tokens"][-1], b["score"]) for b in cands])
        beams = cands[:beam_size]                                              # :280
        if return_trace or return_cands:
            trace.append([(b["parent"], b["tokens"][-1], b["score"]) for b in beams])
        if all(b["tokens"][-1] == end for b in beams):                         # :282-284
            completed.extend(beams)
            break
    best = max(completed, key=lambda b: b["score"]) if completed else beams[0]  # :286-290
    seq = best["tokens"]
    if seq and seq[0] == start:
        seq = seq[1:]
    if end in seq:
        seq = seq[: seq.index(end)]
    if return_cands:
        return seq, best["score"], trace, all_cands
    if return_trace:
        return seq, best["score"], trace
    return seq


def beam_search_batched(p: Params, encoder_output: torch.Tensor, start: int, end: int,
                        max_length: int, beam_size: int, cfg: dict):
    """CPU spec of batched beam (BASELINE config 3): the reference's B==1 beam
    run independently per image.  (The reference itself falls
    back to greedy for B>1, seq2seq.py:244-247.)"""
    return [beam_search(p, encoder_output[i:i + 1], start, end, max_length, beam_size, cfg)
            for i in range(encoder_output.shape[0])]
