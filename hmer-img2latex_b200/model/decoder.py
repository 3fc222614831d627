"""Drop-in decoder modules: same constructor signatures, attributes and ``state_dict``
keys as the reference's ``img2latex/model/decoder.py`` (LSTMDecoder 16-284, Attention
287-343).  ``decode_step`` / ``Attention.forward`` and the three device-resident
decode loops run through the C-ABI; the torch sub-modules only own the parameters.
"""
from __future__ import annotations

import ctypes as C
from typing import Optional, Tuple

import torch
import torch.nn as nn

from .. import _native as N
from .. import ops as _ops  # noqa: F401  (registers torch.ops.i2l.*)
from ._common import Workspace, default_precision, f32c, params_key, require_cuda


class Attention(nn.Module):
    """reference: img2latex/model/decoder.py:287-343."""

    def __init__(self, hidden_dim: int, encoder_dim: int):
        super().__init__()
        self.hidden_dim, self.encoder_dim = hidden_dim, encoder_dim
        self.attn = nn.Linear(hidden_dim + encoder_dim, hidden_dim)
        self.v = nn.Linear(hidden_dim, 1, bias=False)
        self._ws = Workspace()

    def forward(self, hidden: torch.Tensor, encoder_outputs: torch.Tensor) -> torch.Tensor:
        """hidden (B,1,H), encoder_outputs (B,L,E) -> context (B,1,E); decoder.py:312-343."""
        require_cuda(hidden, "Attention.forward")
        if hidden.dim() != 3 or hidden.shape[1] != 1 or encoder_outputs.dim() != 3:
            raise RuntimeError(f"Attention expects hidden (B,1,H) and encoder_outputs (B,L,E); got "
                               f"{tuple(hidden.shape)}, {tuple(encoder_outputs.shape)}")
        B, L, E = encoder_outputs.shape
        with torch.cuda.device(hidden.device):
            lib = N.lib()
            hid, enc = f32c(hidden.reshape(B, self.hidden_dim)), f32c(encoder_outputs)
            w, b, v = f32c(self.attn.weight), f32c(self.attn.bias), f32c(self.v.weight)
            wsb = lib.i2l_attention_workspace_bytes(self.hidden_dim, E, B, L)
            ws = self._ws.get(wsb, hidden.device)
            out = torch.ops.i2l.attention_fwd(hid, enc, w, b, v, ws)
        return out.unsqueeze(1)


class LSTMDecoder(nn.Module):
    """reference: img2latex/model/decoder.py:16-284."""

    def __init__(self, vocab_size: int, embedding_dim: int = None, hidden_dim: int = None,
                 max_seq_length: int = None, lstm_layers: int = None, dropout: float = None,
                 attention: bool = True, precision: Optional[str] = None):
        super().__init__()
        # defaults: decoder.py:48-58
        embedding_dim = 256 if embedding_dim is None else embedding_dim
        hidden_dim = 256 if hidden_dim is None else hidden_dim
        max_seq_length = 141 if max_seq_length is None else max_seq_length
        lstm_layers = 1 if lstm_layers is None else lstm_layers
        dropout = 0.1 if dropout is None else dropout
        if lstm_layers > N.MAX_LSTM:
            raise ValueError(f"at most {N.MAX_LSTM} LSTM layers are supported")
        self.vocab_size, self.embedding_dim, self.hidden_dim = vocab_size, embedding_dim, hidden_dim
        self.max_seq_length, self.lstm_layers, self.dropout = max_seq_length, lstm_layers, dropout
        self.use_attention = attention
        self.precision = precision or default_precision()
        # bf16 only: run the stream-ordered path (one group of launches per step) even where a persistent cluster
        # kernel exists (I2L_BF16_STREAMED) -- the A/B reference for the persistent kernels, same packed weights
        self.streamed = False
        self.embedding = nn.Embedding(vocab_size, embedding_dim)
        self.lstm = nn.LSTM(input_size=2 * embedding_dim, hidden_size=hidden_dim, num_layers=lstm_layers,
                            batch_first=True, dropout=dropout if lstm_layers > 1 else 0)
        if attention:
            self.attention = Attention(hidden_dim, embedding_dim)
        self.output_layer = nn.Linear(hidden_dim, vocab_size)
        self.dropout_layer = nn.Dropout(dropout)
        self._packed = None
        self._packed_key = None
        self._ws = Workspace()

    # -- native plumbing -------------------------------------------------
    def _desc(self) -> N.DecDesc:
        d = N.DecDesc()
        d.vocab_size, d.embedding_dim, d.hidden_dim = self.vocab_size, self.embedding_dim, self.hidden_dim
        d.lstm_layers, d.attention = self.lstm_layers, int(self.use_attention)
        d.precision = N.PRECISIONS[self.precision]
        if self.streamed and self.precision == "bf16":
            d.precision = N.BF16_STREAMED
        return d

    def _desc_list(self):
        """the descriptor as the int list the custom ops take (ops.py)."""
        d = self._desc()
        return [d.vocab_size, d.embedding_dim, d.hidden_dim, d.lstm_layers, d.attention, d.precision]

    def _weights(self):
        ts = [self.embedding.weight]
        for l in range(self.lstm_layers):
            ts += [getattr(self.lstm, f"weight_ih_l{l}"), getattr(self.lstm, f"weight_hh_l{l}"),
                   getattr(self.lstm, f"bias_ih_l{l}"), getattr(self.lstm, f"bias_hh_l{l}")]
        return ts + [self.output_layer.weight, self.output_layer.bias]

    def _ensure_packed(self, device):
        ts = self._weights()
        key = params_key(ts, self.precision)
        if self._packed is not None and key == self._packed_key:
            return
        lib = N.lib()
        d = self._desc()
        held = [f32c(t) for t in ts]
        p = N.DecParams()
        p.embedding = held[0].data_ptr()
        for l in range(self.lstm_layers):
            p.w_ih[l], p.w_hh[l], p.b_ih[l], p.b_hh[l] = (held[1 + 4 * l + j].data_ptr() for j in range(4))
        p.out_w, p.out_b = held[-2].data_ptr(), held[-1].data_ptr()
        nbytes = lib.i2l_dec_packed_bytes(C.byref(d))
        packed = torch.empty(nbytes, dtype=torch.uint8, device=device)
        N.check(lib.i2l_dec_pack(C.byref(d), C.byref(p), N.ptr(packed), nbytes, N.stream_ptr(device)), "i2l_dec_pack")
        torch.cuda.current_stream(device).synchronize()
        self._packed, self._packed_key = packed, key

    def _workspace(self, rows: int, max_length: int, device) -> torch.Tensor:
        d = self._desc()
        return self._ws.get(N.lib().i2l_dec_workspace_bytes(C.byref(d), rows, max_length), device)

    # -- reference API ---------------------------------------------------
    def forward(self, encoder_output: torch.Tensor, target_sequence: torch.Tensor, hidden=None,
                return_hidden: bool = False):
        """Teacher-forced pass, reference decoder.py:100-195 in eval mode (dropout is the identity; the
        backward pass stays out of scope): (B,E), (B,T) int64 -> logits (B,T,V).  ``return_hidden=True``
        additionally returns the final (h, c)."""
        require_cuda(encoder_output, "LSTMDecoder.forward")
        if self.training and self.dropout > 0:
            raise RuntimeError("LSTMDecoder.forward: only the eval-mode forward runs on the native path "
                               "(call .eval(); training with dropout / autograd is out of scope)")
        if target_sequence.dim() != 2:
            raise ValueError(f"target_sequence must be (B,T), got {tuple(target_sequence.shape)}")
        dev = encoder_output.device
        with torch.cuda.device(dev):
            self._ensure_packed(dev)
            lib, d = N.lib(), self._desc()
            B, T = target_sequence.shape
            enc = f32c(encoder_output)
            tgt = target_sequence.to(device=dev, dtype=torch.int64).contiguous()
            h_in = c_in = None
            if hidden is not None:
                h_in, c_in = f32c(hidden[0]), f32c(hidden[1])
            wsb = lib.i2l_dec_forward_workspace_bytes(C.byref(d), max(B, 1), max(T, 1))
            ws = self._ws.get(wsb, dev)
            logits, h_out, c_out, bad = torch.ops.i2l.decoder_forward(enc, tgt, h_in, c_in, self._packed, ws,
                                                                      self._desc_list())
            if int(bad.item()):                                      # nn.Embedding: IndexError
                raise IndexError("index out of range in self (token id outside [0, vocab_size))")
        if return_hidden:
            return logits, (h_out, c_out)
        return logits

    def decode_step(self, encoder_output: torch.Tensor, input_token: torch.Tensor, hidden=None
                    ) -> Tuple[torch.Tensor, Tuple[torch.Tensor, torch.Tensor]]:
        """(B,E), (B,1) int64, None | ((L,B,H),(L,B,H)) -> logits (B,1,V), (h,c);
        reference decoder.py:197-284."""
        require_cuda(encoder_output, "LSTMDecoder.decode_step")
        if input_token.dim() != 2 or input_token.shape[1] != 1:
            raise RuntimeError(f"Shape mismatch: input_token must be (B,1), got {tuple(input_token.shape)}")
        dev = encoder_output.device
        with torch.cuda.device(dev):
            self._ensure_packed(dev)
            lib = N.lib()
            d = self._desc()
            B = input_token.shape[0]
            enc = f32c(encoder_output)
            tok = input_token.reshape(B).to(torch.int64).contiguous()
            h_in = c_in = None
            if hidden is not None:
                h_in, c_in = f32c(hidden[0]), f32c(hidden[1])
            ws = self._workspace(max(B, 1), 1, dev)
            logits, h_out, c_out, bad = torch.ops.i2l.decode_step(enc, tok, h_in, c_in, self._packed, ws,
                                                                  self._desc_list())
            if int(bad.item()):                                      # nn.Embedding: IndexError (decoder.py:232)
                raise IndexError("index out of range in self (token id outside [0, vocab_size))")
        return logits.unsqueeze(1), (h_out, c_out)

    # -- device-resident loops (no host sync per token) ------------------
    def greedy(self, encoder_output: torch.Tensor, start_token_id: int, end_token_id: int, max_length: int,
               temperature: float = 1.0, stop_rule: int = N.STOP_ALL_END_SAME_STEP):
        """Loop of Seq2SeqModel._greedy_search (seq2seq.py:192-232) on the device.
        Returns tokens (B, max_length+1) int64, lengths (B) int32, steps_run (0-dim int32)."""
        require_cuda(encoder_output, "LSTMDecoder.greedy")
        dev = encoder_output.device
        with torch.cuda.device(dev):
            self._ensure_packed(dev)
            lib, d = N.lib(), self._desc()
            enc = f32c(encoder_output)
            B = enc.shape[0]
            ws = self._workspace(max(B, 1), max_length, dev)
            return torch.ops.i2l.decode_greedy(enc, self._packed, ws, self._desc_list(), int(start_token_id),
                                               int(end_token_id), int(max_length), float(temperature), int(stop_rule))

    def sample(self, encoder_output: torch.Tensor, start_token_id: int, end_token_id: int, max_length: int,
               temperature: float = 1.0, top_k: int = 0, top_p: float = 0.0, seed: int = 0, offset: int = 0,
               uniforms: Optional[torch.Tensor] = None, return_probs: bool = False):
        """Loop of Predictor.predict_batch (predictor.py:283-347) on the device."""
        require_cuda(encoder_output, "LSTMDecoder.sample")
        dev = encoder_output.device
        with torch.cuda.device(dev):
            self._ensure_packed(dev)
            lib, d = N.lib(), self._desc()
            enc = f32c(encoder_output)
            B = enc.shape[0]
            if uniforms is not None:
                uniforms = f32c(uniforms.to(dev))
                if tuple(uniforms.shape) != (max_length, B):
                    raise RuntimeError(f"uniforms must be (max_length, B) = ({max_length}, {B})")
            ws = self._workspace(max(B, 1), max_length, dev)
            tokens, lengths, steps, probs = torch.ops.i2l.decode_sample(
                enc, self._packed, ws, self._desc_list(), int(start_token_id), int(end_token_id), int(max_length),
                float(temperature), int(top_k), float(top_p), int(seed), int(offset), uniforms, bool(return_probs))
        if return_probs:
            return tokens, lengths, steps, probs
        return tokens, lengths, steps

    def beam(self, encoder_output: torch.Tensor, start_token_id: int, end_token_id: int, max_length: int,
             beam_size: int, return_trace: bool = False, return_candidates: bool = False):
        """Seq2SeqModel._beam_search (seq2seq.py:234-298) run independently per image, on the
        device.  Returns out_tokens (B,max_length) int64 padded with -1, out_len (B), score (B) f64
        [, (parent, token, score) traces (T,B,K)] [, (token, log-prob) candidate audit trail (T,B,K,K)]."""
        require_cuda(encoder_output, "LSTMDecoder.beam")
        dev = encoder_output.device
        with torch.cuda.device(dev):
            self._ensure_packed(dev)
            lib, d = N.lib(), self._desc()
            enc = f32c(encoder_output)
            B, K = enc.shape[0], beam_size
            ws = self._workspace(max(B * K, 1), max_length, dev)
            out, olen, score, trp, trt, trs, ctok, clogp = torch.ops.i2l.decode_beam(
                enc, self._packed, ws, self._desc_list(), int(K), int(start_token_id), int(end_token_id), int(max_length),
                bool(return_trace), bool(return_candidates))
        res = (out, olen, score)
        if return_trace:
            res += ((trp, trt, trs),)
        if return_candidates:      # (T,B,K,K) audit trail: every live beam's top-K (token, log-prob), see i2l_decode_beam
            res += ((ctok, clogp),)
        return res
