"""Drop-in ``Seq2SeqModel`` (reference: img2latex/model/seq2seq.py:17-298): same
constructor, sub-module names (``encoder`` / ``decoder`` => same ``state_dict`` keys),
``inference`` signature and return conventions.  The decode loops run on the device
with no host sync per token; the host reads the token matrix back once at the end.
"""
from __future__ import annotations

from typing import Dict, Iterable, Iterator, List, Optional, Tuple

import torch
import torch.nn as nn

from .. import _native as N
from .decoder import LSTMDecoder
from .encoder import CNNEncoder, ResNetEncoder, normalize_u8


def _checked_steps(n: int) -> int:
    """steps_run < 0 is the device-side report of a decode loop that aborted (decode_wide.cu: a bounded spin expired)"""
    if n < 0:
        raise RuntimeError("the persistent decode loop aborted on the device (a bounded flag wait expired); the tokens "
                           "of this call are invalid")
    return n


class Seq2SeqModel(nn.Module):
    def __init__(self, model_type: str = "cnn_lstm", vocab_size: int = None, encoder_params: Dict = None,
                 decoder_params: Dict = None, precision: Optional[str] = None):
        super().__init__()
        encoder_params = {} if encoder_params is None else encoder_params
        decoder_params = {} if decoder_params is None else decoder_params
        vocab_size = 100 if vocab_size is None else vocab_size              # seq2seq.py:50-51
        embedding_dim = encoder_params.get("embedding_dim", 256)            # seq2seq.py:54
        if model_type == "cnn_lstm":                                        # seq2seq.py:57-67
            self.encoder = CNNEncoder(
                img_height=encoder_params.get("img_height", 50), img_width=encoder_params.get("img_width", 200),
                channels=encoder_params.get("channels", 1),
                conv_filters=encoder_params.get("conv_filters", [32, 64, 128]),
                kernel_size=encoder_params.get("kernel_size", 3), pool_size=encoder_params.get("pool_size", 2),
                padding=encoder_params.get("padding", "same"), embedding_dim=embedding_dim, precision=precision)
        elif model_type == "resnet_lstm":                                   # seq2seq.py:68-76
            self.encoder = ResNetEncoder(
                img_height=encoder_params.get("img_height", 224), img_width=encoder_params.get("img_width", 224),
                channels=encoder_params.get("channels", 3), model_name=encoder_params.get("model_name", "resnet50"),
                embedding_dim=embedding_dim, freeze_backbone=encoder_params.get("freeze_backbone", True),
                pretrained=encoder_params.get("pretrained", False), precision=precision)
        else:
            raise ValueError(f"Invalid model type: {model_type}. Expected 'cnn_lstm' or 'resnet_lstm'.")
        self.decoder = LSTMDecoder(                                         # seq2seq.py:83-91
            vocab_size=vocab_size, embedding_dim=embedding_dim, hidden_dim=decoder_params.get("hidden_dim", 256),
            max_seq_length=decoder_params.get("max_seq_length", 150), lstm_layers=decoder_params.get("lstm_layers", 1),
            dropout=decoder_params.get("dropout", 0.1), attention=decoder_params.get("attention", False),
            precision=precision)
        self.model_type = model_type
        self.vocab_size = vocab_size

    def set_precision(self, precision: str) -> "Seq2SeqModel":
        if precision not in N.PRECISIONS:
            raise ValueError(f"precision must be one of {sorted(N.PRECISIONS)}")
        self.encoder.precision = precision
        self.decoder.precision = precision
        return self

    def forward(self, images: torch.Tensor, target_sequences: torch.Tensor) -> torch.Tensor:
        """reference seq2seq.py:98-122 in eval mode: encoder, then the teacher-forced decoder pass over
        ``target_sequences[:, :-1]`` -> logits (B, T-1, V) (validation loss / accuracy, trainer.py:508-533)."""
        encoder_output = self.encoder(images)
        return self.decoder(encoder_output, target_sequences[:, :-1])

    @torch.no_grad()
    def inference(self, image: torch.Tensor, start_token_id: int, end_token_id: int, max_length: int = None,
                  temperature: float = None, top_k: int = None, top_p: float = None, beam_size: int = None):
        """reference seq2seq.py:124-190.  B==1 -> List[int] (START stripped, cut at END);
        B>1 -> raw List[List[int]] incl. START (seq2seq.py:223-232)."""
        max_length = 150 if max_length is None else max_length
        temperature = 1.0 if temperature is None else temperature
        top_k = 0 if top_k is None else top_k
        top_p = 0.0 if top_p is None else top_p
        beam_size = 0 if beam_size is None else beam_size
        encoder_output = self.encoder(image)
        if encoder_output.dim() == 1:
            encoder_output = encoder_output.unsqueeze(0)
        if beam_size > 0:
            return self._beam_search(encoder_output, start_token_id, end_token_id, max_length, beam_size)
        return self._greedy_search(encoder_output, start_token_id, end_token_id, max_length, temperature, top_k, top_p)

    def _greedy_search(self, encoder_output, start_token_id, end_token_id, max_length, temperature, top_k, top_p):
        """reference seq2seq.py:192-232 (top_k / top_p are ignored there too)."""
        tokens, _, steps = self.decoder.greedy(encoder_output, start_token_id, end_token_id, max_length,
                                               temperature, N.STOP_ALL_END_SAME_STEP)
        n = _checked_steps(int(steps.item()))       # the one host sync of the decode
        rows = tokens[:, : n + 1].tolist()
        if len(rows) == 1:                          # seq2seq.py:224-231
            seq = rows[0]
            if seq and seq[0] == start_token_id:
                seq = seq[1:]
            if end_token_id in seq:
                seq = seq[: seq.index(end_token_id)]
            return seq
        return rows

    def _beam_search(self, encoder_output, start_token_id, end_token_id, max_length, beam_size):
        """reference seq2seq.py:234-298: batch-size-1 only, B>1 falls back to greedy (244-247)."""
        if encoder_output.size(0) != 1:
            return self._greedy_search(encoder_output, start_token_id, end_token_id, max_length, 1.0, 0, 0.0)
        return self.beam_search_batch(encoder_output, start_token_id, end_token_id, max_length, beam_size)[0]

    @torch.no_grad()
    def beam_search_batch(self, encoder_output, start_token_id, end_token_id, max_length, beam_size) -> List[List[int]]:
        """Batched beam search = the reference's B==1 beam run independently per image
        (BASELINE config 3); not present in the reference."""
        out, olen, _ = self.decoder.beam(encoder_output, start_token_id, end_token_id, max_length, beam_size)
        out, olen = out.tolist(), olen.tolist()
        return [r[:n] for r, n in zip(out, olen)]

    @torch.no_grad()
    def greedy_stream(self, host_batches: Iterable[torch.Tensor], start_token_id: int, end_token_id: int,
                      max_length: int = 150, temperature: float = 1.0, stop_rule: int = N.STOP_ALL_END_SAME_STEP,
                      device: Optional[torch.device] = None, normalize: str = "pm1", exchange=None,
                      exchange_readback: str = "global") -> Iterator[Tuple[torch.Tensor, torch.Tensor, int]]:
        """Serving loop over HOST batches (ideally pinned (B,C,H,W) tensors: fp32 like the
        reference's, bf16, or raw uint8 pixels that are normalised on the device with
        ``normalize`` = "pm1" | "meanstd", see ``normalize_u8``): the host->device
        copy of batch i+1 runs on a copy stream while batch i is encoded and decoded, and the
        token ids come back through a pinned buffer.  Yields (tokens (B,max_length+1) int64 on the
        host, lengths (B) int32 on the host, steps_run) per batch -- same content as
        encoder + LSTMDecoder.greedy.  The yielded tensors are pinned buffers owned by the model (two slots, reused by
        later batches AND later calls): copy what must outlive the next iteration.  Not in the reference (its Predictor copies and computes
        serially, training/predictor.py:245-262).

        ``exchange`` (a ``dist.TokenExchange``, batch-sharded multi-GPU serving): every rank streams ITS shard of each
        global batch; the ids are exchanged on the device by direct peer stores and the yielded triple is the GLOBAL
        (n_total, max_length+1) result -- two batches later than without (the exchange of batch i runs on a side
        stream beside the decode kernel of batch i + 1 and is handed out during batch i + 2), all global batches having
        the same size.  ``exchange_readback="shard"``: every rank still holds the
        global ids in HBM after the exchange, but copies only its OWN rows to the host (the rank that hands the job's
        result to the consumer uses "global"; N host copies of the same matrix are N - 1 too many)."""
        dev = device or next(self.parameters()).device
        if dev.type != "cuda":
            raise RuntimeError("greedy_stream needs the model on a CUDA device")
        copy_stream = torch.cuda.Stream(dev)
        out_stream = torch.cuda.Stream(dev)                       # device->host copies of the results: off the compute stream
        compute = torch.cuda.current_stream(dev)
        computed = [torch.cuda.Event(), torch.cuda.Event()]       # results of the slot exist on the device

        def read_back(slot: int, tokens, lengths, steps, after=None) -> None:
            if out_tok[slot] is None or out_tok[slot].shape != tokens.shape:
                # pinned result buffers are kept on the model across calls: cudaHostAlloc costs milliseconds, and a
                # short stream would pay it for every slot inside its own run
                key = (slot, tuple(tokens.shape), tuple(lengths.shape), tokens.dtype, lengths.dtype, steps.dtype)
                cache = self.__dict__.setdefault("_pinned_results", {})
                if key not in cache:
                    if len(cache) >= 8:
                        cache.clear()
                    cache[key] = (torch.empty(tokens.shape, dtype=tokens.dtype).pin_memory(),
                                  torch.empty(lengths.shape, dtype=lengths.dtype).pin_memory(),
                                  torch.empty((), dtype=steps.dtype).pin_memory())
                out_tok[slot], out_len[slot], out_steps[slot] = cache[key]
            if after is None:
                computed[slot].record(compute)
            with torch.cuda.stream(out_stream):
                out_stream.wait_event(computed[slot] if after is None else after)
                for dst, src in ((out_tok[slot], tokens), (out_len[slot], lengths), (out_steps[slot], steps)):
                    dst.copy_(src, non_blocking=True)
                    src.record_stream(out_stream)                 # the caching allocator must not recycle it under the copy
                done[slot].record(out_stream)

        NS = 3                                                    # input staging slots: two copies stay queued ahead of the compute
        bufs: List[Optional[torch.Tensor]] = [None] * NS
        ready = [torch.cuda.Event() for _ in range(NS)]           # H2D copy of the slot has landed
        consumed = [torch.cuda.Event() for _ in range(NS)]        # the encoder has read the slot
        done = [torch.cuda.Event(), torch.cuda.Event()]           # results of the (output) slot are in host memory
        used = [False] * NS
        out_tok: List[Optional[torch.Tensor]] = [None, None]
        out_len: List[Optional[torch.Tensor]] = [None, None]
        out_steps: List[Optional[torch.Tensor]] = [None, None]

        def stage(slot: int, xb: torch.Tensor) -> None:
            with torch.cuda.stream(copy_stream):
                if bufs[slot] is None or bufs[slot].shape != xb.shape or bufs[slot].dtype != xb.dtype:
                    # (re)allocate ON the copy stream, ordered after everything queued on the compute stream: the
                    # caching allocator may hand back a block whose last kernels (an encoder temporary, the caller's
                    # own work) are still in flight there; record_stream keeps the block alive for the reader
                    copy_stream.wait_stream(compute)
                    bufs[slot] = torch.empty(xb.shape, dtype=xb.dtype, device=dev)
                    bufs[slot].record_stream(compute)
                if used[slot]:
                    copy_stream.wait_event(consumed[slot])        # the batch that used this buffer has been encoded
                bufs[slot].copy_(xb, non_blocking=True)
                ready[slot].record(copy_stream)

        def collect(slot: int):
            done[slot].synchronize()                              # the one host sync of the batch
            return out_tok[slot], out_len[slot], _checked_steps(int(out_steps[slot]))

        # Three batches are in flight: the copies of batches i+1 and i+2 are queued while batch i computes (a host hiccup
        # of up to a whole batch time then leaves the copy engine busy), and the kernels of batch i+1 are enqueued BEFORE
        # the host waits for the results of batch i, so the GPU never idles on the host.  A yielded triple lives in
        # pinned buffers that are reused: it is valid until the generator is advanced.
        it = iter(host_batches)
        queue: List[int] = []                                     # staged input slots, oldest first
        n_in = 0

        def stage_next() -> None:
            nonlocal n_in
            xb = next(it, None)
            if xb is not None:
                stage(n_in % NS, xb)
                queue.append(n_in % NS)
                n_in += 1

        stage_next()
        stage_next()
        if not queue:
            return
        prev = -1
        nres = 0

        def shard_view(res):
            tokens, lengths, steps = res
            if exchange is not None and exchange_readback == "shard":
                lo = exchange.shard_lo
                tokens, lengths = tokens[lo: lo + exchange.shard], lengths[lo: lo + exchange.shard]
            return tokens, lengths, steps

        while queue:
            slot = queue.pop(0)
            stage_next()                                          # overlaps with the compute below
            compute.wait_event(ready[slot])
            xin = bufs[slot]
            if xin.dtype == torch.uint8 and hasattr(self.encoder, "forward_u8"):
                enc = self.encoder.forward_u8(xin, normalize)     # normalisation fused into conv1 where supported
            else:
                if xin.dtype == torch.uint8:
                    xin = normalize_u8(xin, normalize, out_dtype=torch.bfloat16 if self.encoder.precision == "bf16"
                                       else torch.float32)
                enc = self.encoder(xin)
            consumed[slot].record(compute)
            used[slot] = True
            res, ev = None, None
            if exchange is not None:
                # release the exchange of the PREVIOUS batch's shard result here, between this batch's encoder and its
                # decode (it then runs beside the decode kernel); hands out the global result of the batch before that
                res, ev = exchange.kick(), exchange.result_event
            tokens, lengths, steps = self.decoder.greedy(enc, start_token_id, end_token_id, max_length, temperature,
                                                         stop_rule)
            if exchange is not None:
                exchange.stage(tokens, lengths, steps)
            else:
                res = (tokens, lengths, steps)
            if res is not None:
                oslot = nres & 1
                nres += 1
                # exchanged results are produced on the exchange's side stream: the copy waits for THAT event
                read_back(oslot, *shard_view(res), after=ev)
                if prev >= 0:
                    yield collect(prev)
                prev = oslot
        if exchange is not None:
            for res in exchange.flush():                          # the last two batches' global results
                oslot = nres & 1
                nres += 1
                read_back(oslot, *shard_view(res))
                if prev >= 0:
                    yield collect(prev)
                prev = oslot
        if prev >= 0:
            yield collect(prev)
