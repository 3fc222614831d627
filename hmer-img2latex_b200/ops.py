"""PyTorch custom operators (``torch.library``, namespace ``i2l``) over the C-ABI of ``libi2l_b200.so``.

BASELINE.json's north_star: "the reference's encoder/decoder modules and decode strategies can be swapped for
PyTorch custom ops that call hand-written sm_100a CUDA kernels through a thin C-ABI layer".  This module is that
layer: every op is registered for the CUDA dispatch key only (there is NO CPU kernel -- calling an op with CPU
tensors fails in the dispatcher), carries a fake (meta) kernel so that shapes / dtypes propagate under
``torch.compile`` / ``torch.export`` graph capture, and does nothing but marshal tensors into the ``extern "C"``
call on the current CUDA stream.  The drop-in modules in ``model/`` call ``torch.ops.i2l.*`` and nothing else.

op                         reference symbol it replaces (file:line under /root/reference/img2latex)
i2l::cnn_encoder_fwd       CNNEncoder.forward                 model/encoder.py:111-129
i2l::cnn_encoder_fwd_u8    load_image / _prepare_image pixel arithmetic + CNNEncoder.forward
                                                              data/utils.py:68-80, training/predictor.py:441-446
i2l::resnet_encoder_fwd    ResNetEncoder.forward              model/encoder.py:231-249
i2l::attention_fwd         Attention.forward                  model/decoder.py:312-343
i2l::decode_step           LSTMDecoder.decode_step            model/decoder.py:197-284
i2l::decoder_forward       LSTMDecoder.forward (eval)         model/decoder.py:100-195
i2l::decode_greedy         Seq2SeqModel._greedy_search loop   model/seq2seq.py:192-232
i2l::decode_sample         Predictor.predict_batch loop       training/predictor.py:283-347
i2l::decode_beam           Seq2SeqModel._beam_search          model/seq2seq.py:234-298

Descriptors travel as int lists: ``cnn_desc = [img_height, img_width, channels, kernel_size, pool_size,
embedding_dim, precision, *conv_filters]``, ``resnet_desc = [depth, img_height, embedding_dim, precision]``,
``dec_desc = [vocab_size, embedding_dim, hidden_dim, lstm_layers, attention, precision]``.  ``packed`` is the
uint8 buffer written by ``i2l_*_pack``; ``workspace`` is a caller-owned uint8 scratch buffer (declared as mutated).
"""
from __future__ import annotations

import ctypes as C
from typing import List, Optional, Tuple

import torch
from torch import Tensor

from . import _native as N

_T = torch.library.custom_op


def _cnn_desc(v: List[int]) -> N.CnnDesc:
    d = N.CnnDesc()
    d.img_height, d.img_width, d.channels, d.kernel_size, d.pool_size, d.embedding_dim, d.precision = v[:7]
    d.n_conv = len(v) - 7
    for i, f in enumerate(v[7:]):
        d.filters[i] = f
    return d


def _resnet_desc(v: List[int]) -> N.ResnetDesc:
    d = N.ResnetDesc()
    d.depth, d.img_height, d.embedding_dim, d.precision = v
    return d


def _dec_desc(v: List[int]) -> N.DecDesc:
    d = N.DecDesc()
    d.vocab_size, d.embedding_dim, d.hidden_dim, d.lstm_layers, d.attention, d.precision = v
    return d


def _sp(t: Tensor) -> C.c_void_p:
    return N.stream_ptr(t.device)


# ------------------------------------------------------------------------------------------------ encoders
@_T("i2l::cnn_encoder_fwd", mutates_args=("workspace",), device_types="cuda")
def cnn_encoder_fwd(x: Tensor, packed: Tensor, workspace: Tensor, cnn_desc: List[int]) -> Tensor:
    d = _cnn_desc(cnn_desc)
    B = x.shape[0]
    out = torch.empty(B, d.embedding_dim, dtype=torch.float32, device=x.device)
    if B:
        in_dtype = N.IN_BF16 if x.dtype == torch.bfloat16 else N.IN_F32
        with torch.cuda.device(x.device):
            N.check(N.lib().i2l_cnn_encoder_fwd_in(C.byref(d), N.ptr(packed), N.ptr(x), in_dtype, B, N.ptr(out),
                                                   N.ptr(workspace), workspace.numel(), _sp(x)), "i2l_cnn_encoder_fwd")
    return out


@cnn_encoder_fwd.register_fake
def _(x, packed, workspace, cnn_desc):
    return x.new_empty((x.shape[0], cnn_desc[5]), dtype=torch.float32)


@_T("i2l::cnn_encoder_fwd_u8", mutates_args=("workspace",), device_types="cuda")
def cnn_encoder_fwd_u8(pixels: Tensor, packed: Tensor, workspace: Tensor, cnn_desc: List[int], norm_mode: int,
                       mean: List[float], std: List[float]) -> Tensor:
    d = _cnn_desc(cnn_desc)
    B = pixels.shape[0]
    out = torch.empty(B, d.embedding_dim, dtype=torch.float32, device=pixels.device)
    if B:
        m = (C.c_float * 4)(*([float(v) for v in mean[:3]] + [0.0]))
        sd = (C.c_float * 4)(*([float(v) for v in std[:3]] + [1.0]))
        with torch.cuda.device(pixels.device):
            N.check(N.lib().i2l_cnn_encoder_fwd_u8(C.byref(d), N.ptr(packed), N.ptr(pixels), norm_mode, m, sd, B,
                                                   N.ptr(out), N.ptr(workspace), workspace.numel(), _sp(pixels)),
                    "i2l_cnn_encoder_fwd_u8")
    return out


@cnn_encoder_fwd_u8.register_fake
def _(pixels, packed, workspace, cnn_desc, norm_mode, mean, std):
    return pixels.new_empty((pixels.shape[0], cnn_desc[5]), dtype=torch.float32)


@_T("i2l::resnet_encoder_fwd", mutates_args=("workspace",), device_types="cuda")
def resnet_encoder_fwd(x: Tensor, packed: Tensor, workspace: Tensor, resnet_desc: List[int]) -> Tensor:
    d = _resnet_desc(resnet_desc)
    B, W = x.shape[0], x.shape[3]
    out = torch.empty(B, d.embedding_dim, dtype=torch.float32, device=x.device)
    if B:
        with torch.cuda.device(x.device):
            N.check(N.lib().i2l_resnet_encoder_fwd(C.byref(d), N.ptr(packed), N.ptr(x), B, W, N.ptr(out),
                                                   N.ptr(workspace), workspace.numel(), _sp(x)), "i2l_resnet_encoder_fwd")
    return out


@resnet_encoder_fwd.register_fake
def _(x, packed, workspace, resnet_desc):
    return x.new_empty((x.shape[0], resnet_desc[2]), dtype=torch.float32)


# ------------------------------------------------------------------------------------------------ decoder
@_T("i2l::attention_fwd", mutates_args=("workspace",), device_types="cuda")
def attention_fwd(hidden: Tensor, encoder_outputs: Tensor, attn_w: Tensor, attn_b: Tensor, v_w: Tensor,
                  workspace: Tensor) -> Tensor:
    B, L, E = encoder_outputs.shape
    H = hidden.shape[-1]
    out = torch.empty(B, E, dtype=torch.float32, device=hidden.device)
    with torch.cuda.device(hidden.device):
        N.check(N.lib().i2l_attention_fwd(H, E, N.ptr(attn_w), N.ptr(attn_b), N.ptr(v_w), N.ptr(hidden),
                                          N.ptr(encoder_outputs), B, L, N.ptr(out), N.ptr(workspace), workspace.numel(),
                                          _sp(hidden)), "i2l_attention_fwd")
    return out


@attention_fwd.register_fake
def _(hidden, encoder_outputs, attn_w, attn_b, v_w, workspace):
    return hidden.new_empty((encoder_outputs.shape[0], encoder_outputs.shape[2]), dtype=torch.float32)


@_T("i2l::decode_step", mutates_args=("workspace",), device_types="cuda")
def decode_step(enc: Tensor, tok: Tensor, h_in: Optional[Tensor], c_in: Optional[Tensor], packed: Tensor,
                workspace: Tensor, dec_desc: List[int]) -> Tuple[Tensor, Tensor, Tensor, Tensor]:
    """-> logits (B,V), h (L,B,H), c (L,B,H), bad (int32 scalar: 1 when an id was outside [0,V))."""
    d = _dec_desc(dec_desc)
    B, dev = tok.shape[0], enc.device
    logits = torch.empty(B, d.vocab_size, dtype=torch.float32, device=dev)
    h_out = torch.empty(d.lstm_layers, B, d.hidden_dim, dtype=torch.float32, device=dev)
    c_out = torch.empty_like(h_out)
    bad = torch.zeros((), dtype=torch.int32, device=dev)
    with torch.cuda.device(dev):
        N.check(N.lib().i2l_decode_step(C.byref(d), N.ptr(packed), N.ptr(enc), N.ptr(tok), B, N.ptr(h_in), N.ptr(c_in),
                                        N.ptr(logits), N.ptr(h_out), N.ptr(c_out), N.ptr(bad), N.ptr(workspace),
                                        workspace.numel(), _sp(enc)), "i2l_decode_step")
    return logits, h_out, c_out, bad


@decode_step.register_fake
def _(enc, tok, h_in, c_in, packed, workspace, dec_desc):
    V, _, H, L = dec_desc[:4]
    B = tok.shape[0]
    f = lambda *s: enc.new_empty(s, dtype=torch.float32)
    return f(B, V), f(L, B, H), f(L, B, H), enc.new_empty((), dtype=torch.int32)


@_T("i2l::decoder_forward", mutates_args=("workspace",), device_types="cuda")
def decoder_forward(enc: Tensor, target: Tensor, h_in: Optional[Tensor], c_in: Optional[Tensor], packed: Tensor,
                    workspace: Tensor, dec_desc: List[int]) -> Tuple[Tensor, Tensor, Tensor, Tensor]:
    d = _dec_desc(dec_desc)
    (B, T), dev = target.shape, enc.device
    logits = torch.empty(B, T, d.vocab_size, dtype=torch.float32, device=dev)
    h_out = torch.empty(d.lstm_layers, B, d.hidden_dim, dtype=torch.float32, device=dev)
    c_out = torch.empty_like(h_out)
    bad = torch.zeros((), dtype=torch.int32, device=dev)
    with torch.cuda.device(dev):
        N.check(N.lib().i2l_decoder_forward(C.byref(d), N.ptr(packed), N.ptr(enc), N.ptr(target), B, T, N.ptr(h_in),
                                            N.ptr(c_in), N.ptr(logits), N.ptr(h_out), N.ptr(c_out), N.ptr(bad),
                                            N.ptr(workspace), workspace.numel(), _sp(enc)), "i2l_decoder_forward")
    return logits, h_out, c_out, bad


@decoder_forward.register_fake
def _(enc, target, h_in, c_in, packed, workspace, dec_desc):
    V, _, H, L = dec_desc[:4]
    B, T = target.shape
    f = lambda *s: enc.new_empty(s, dtype=torch.float32)
    return f(B, T, V), f(L, B, H), f(L, B, H), enc.new_empty((), dtype=torch.int32)


@_T("i2l::decode_greedy", mutates_args=("workspace",), device_types="cuda")
def decode_greedy(enc: Tensor, packed: Tensor, workspace: Tensor, dec_desc: List[int], start_id: int, end_id: int,
                  max_length: int, temperature: float, stop_rule: int) -> Tuple[Tensor, Tensor, Tensor]:
    """-> tokens (B,max_length+1) int64, lengths (B) int32, steps_run () int32; the whole loop runs on the device."""
    d = _dec_desc(dec_desc)
    B, dev = enc.shape[0], enc.device
    tokens = torch.empty(B, max_length + 1, dtype=torch.int64, device=dev)
    lengths = torch.empty(B, dtype=torch.int32, device=dev)
    steps = torch.empty((), dtype=torch.int32, device=dev)      # written by every loop's finalize kernel
    with torch.cuda.device(dev):
        N.check(N.lib().i2l_decode_greedy(C.byref(d), N.ptr(packed), N.ptr(enc), B, start_id, end_id, max_length,
                                          float(temperature), stop_rule, N.ptr(tokens), N.ptr(lengths), N.ptr(steps),
                                          N.ptr(workspace), workspace.numel(), _sp(enc)), "i2l_decode_greedy")
    return tokens, lengths, steps


@decode_greedy.register_fake
def _(enc, packed, workspace, dec_desc, start_id, end_id, max_length, temperature, stop_rule):
    B = enc.shape[0]
    return (enc.new_empty((B, max_length + 1), dtype=torch.int64), enc.new_empty((B,), dtype=torch.int32),
            enc.new_empty((), dtype=torch.int32))


@_T("i2l::decode_sample", mutates_args=("workspace",), device_types="cuda")
def decode_sample(enc: Tensor, packed: Tensor, workspace: Tensor, dec_desc: List[int], start_id: int, end_id: int,
                  max_length: int, temperature: float, top_k: int, top_p: float, seed: int, offset: int,
                  uniforms: Optional[Tensor], return_probs: bool) -> Tuple[Tensor, Tensor, Tensor, Tensor]:
    """-> tokens, lengths, steps_run, probs ((max_length,B,V) filtered distributions, or an empty tensor)."""
    d = _dec_desc(dec_desc)
    B, dev = enc.shape[0], enc.device
    tokens = torch.empty(B, max_length + 1, dtype=torch.int64, device=dev)
    lengths = torch.empty(B, dtype=torch.int32, device=dev)
    steps = torch.empty((), dtype=torch.int32, device=dev)      # written by every loop's finalize kernel
    probs = (torch.zeros(max_length, B, d.vocab_size, dtype=torch.float32, device=dev) if return_probs
             else torch.empty(0, dtype=torch.float32, device=dev))
    with torch.cuda.device(dev):
        N.check(N.lib().i2l_decode_sample(C.byref(d), N.ptr(packed), N.ptr(enc), B, start_id, end_id, max_length,
                                          float(temperature), int(top_k), float(top_p), int(seed), int(offset),
                                          N.ptr(uniforms), N.ptr(tokens), N.ptr(lengths), N.ptr(steps),
                                          N.ptr(probs if return_probs else None), N.ptr(workspace), workspace.numel(),
                                          _sp(enc)), "i2l_decode_sample")
    return tokens, lengths, steps, probs


@decode_sample.register_fake
def _(enc, packed, workspace, dec_desc, start_id, end_id, max_length, temperature, top_k, top_p, seed, offset,
      uniforms, return_probs):
    B = enc.shape[0]
    probs = enc.new_empty((max_length, B, dec_desc[0]) if return_probs else (0,), dtype=torch.float32)
    return (enc.new_empty((B, max_length + 1), dtype=torch.int64), enc.new_empty((B,), dtype=torch.int32),
            enc.new_empty((), dtype=torch.int32), probs)


@_T("i2l::decode_beam", mutates_args=("workspace",), device_types="cuda")
def decode_beam(enc: Tensor, packed: Tensor, workspace: Tensor, dec_desc: List[int], beam_size: int, start_id: int,
                end_id: int, max_length: int, return_trace: bool, return_candidates: bool
                ) -> Tuple[Tensor, Tensor, Tensor, Tensor, Tensor, Tensor, Tensor, Tensor]:
    """-> out_tokens (B,T) int64 (-1 padded), out_len (B) int32, score (B) f64, then the optional
    (T,B,K) traces parent / token / score and the (T,B,K,K) candidate audit trail token / log-prob
    (empty tensors when not requested)."""
    d = _dec_desc(dec_desc)
    B, K, T, dev = enc.shape[0], beam_size, max_length, enc.device
    out = torch.empty(B, T, dtype=torch.int64, device=dev)
    olen = torch.empty(B, dtype=torch.int32, device=dev)
    score = torch.empty(B, dtype=torch.float64, device=dev)
    e = lambda dt: torch.empty(0, dtype=dt, device=dev)
    trp, trt, trs, ctok, clogp = e(torch.int32), e(torch.int32), e(torch.float64), e(torch.int32), e(torch.float32)
    if return_trace:
        trp = torch.empty(T, B, K, dtype=torch.int32, device=dev)
        trt = torch.empty(T, B, K, dtype=torch.int32, device=dev)
        trs = torch.empty(T, B, K, dtype=torch.float64, device=dev)
    if return_candidates:
        ctok = torch.full((T, B, K, K), -1, dtype=torch.int32, device=dev)
        clogp = torch.full((T, B, K, K), float("nan"), dtype=torch.float32, device=dev)
    o = lambda t, on: N.ptr(t if on else None)
    with torch.cuda.device(dev):
        N.check(N.lib().i2l_decode_beam(C.byref(d), N.ptr(packed), N.ptr(enc), B, K, start_id, end_id, T, N.ptr(out),
                                        N.ptr(olen), N.ptr(score), o(trp, return_trace), o(trt, return_trace),
                                        o(trs, return_trace), o(ctok, return_candidates), o(clogp, return_candidates),
                                        N.ptr(workspace), workspace.numel(), _sp(enc)), "i2l_decode_beam")
    return out, olen, score, trp, trt, trs, ctok, clogp


@decode_beam.register_fake
def _(enc, packed, workspace, dec_desc, beam_size, start_id, end_id, max_length, return_trace, return_candidates):
    B, K, T = enc.shape[0], beam_size, max_length
    n = lambda s, dt: enc.new_empty(s, dtype=dt)
    tr = (T, B, K) if return_trace else (0,)
    cd = (T, B, K, K) if return_candidates else (0,)
    return (n((B, T), torch.int64), n((B,), torch.int32), n((B,), torch.float64), n(tr, torch.int32),
            n(tr, torch.int32), n(tr, torch.float64), n(cd, torch.int32), n(cd, torch.float32))


OPS = ("cnn_encoder_fwd", "cnn_encoder_fwd_u8", "resnet_encoder_fwd", "attention_fwd", "decode_step",
       "decoder_forward", "decode_greedy", "decode_sample", "decode_beam")
