"""Serving facade mirroring ``img2latex/training/predictor.py`` (Predictor 28-521) for the
hot path: ``predict`` / ``predict_batch`` on image *tensors* with the reference's loop
semantics (sampling only when temperature > 0 and (top_k > 0 or top_p > 0), sticky
finished flags, cut at the first END, beam clamped to greedy -- predictor.py:163-167,
231-235).  ``_prepare_image`` (396-462) takes the reference's four input kinds: paths go
through the batched device-side ``load_image`` (aspect-preserving LANCZOS resize + pad / crop +
normalisation, ``preprocess.load_images``), PIL images through the device-side plain resize
(Pillow's default bicubic) and x/255*2-1, arrays / tensors of the wrong size are resized bilinearly
as in ``_preprocess_tensor`` (464-499).  Every branch resizes to the MODEL's size (the reference
hard-codes 64x800); only file decoding runs on the host."""
from __future__ import annotations

import logging
from typing import List, Optional, Sequence, Union

import numpy as np
import torch
import torch.nn.functional as F

from . import preprocess as P
from .model.encoder import normalize_u8
from .model.seq2seq import Seq2SeqModel
from .tokenizer import LaTeXTokenizer

logger = logging.getLogger(__name__)


class Predictor:
    def __init__(self, model: Seq2SeqModel, tokenizer: LaTeXTokenizer, device: Optional[torch.device] = None):
        if device is None:
            device = torch.device("cuda", torch.cuda.current_device()) if torch.cuda.is_available() else None
        if device is None or torch.device(device).type != "cuda":
            raise RuntimeError("hmer-img2latex_b200.Predictor needs a CUDA (sm_100a) device; there is no CPU path")
        self.device = torch.device(device)
        self.model = model.to(self.device).eval()                 # predictor.py:50-55
        self.tokenizer = tokenizer

    @classmethod
    def from_checkpoint(cls, checkpoint_path: str, device=None, precision: Optional[str] = None,
                        allow_unsafe_pickle: bool = False) -> "Predictor":
        """Reference checkpoints (layout written by training/trainer.py:207-221) consumed exactly as
        training/predictor.py:61-137 does: tokenizer from ``tokenizer_config``, encoder parameters from
        ``config.model.encoder.{cnn|resnet}``, ``config.model.embedding_dim`` (default 256) for both halves,
        ``model_state_dict`` loaded strictly."""
        # The trainer's checkpoints are plain dicts / lists / tensors (training/trainer.py:207-221): the safe unpickler is
        # enough.  Files that need arbitrary objects are loaded only on explicit request (allow_unsafe_pickle=True).
        try:
            ck = torch.load(checkpoint_path, map_location="cpu", weights_only=True)
        except Exception as e:
            if not allow_unsafe_pickle:
                raise RuntimeError(f"{checkpoint_path}: not loadable with weights_only=True ({e}); pass "
                                   "allow_unsafe_pickle=True if the file is trusted") from e
            ck = torch.load(checkpoint_path, map_location="cpu", weights_only=False)
        config = ck.get("config", {})
        model_config = config.get("model", {})
        model_type = model_config.get("name", "cnn_lstm")
        tok = LaTeXTokenizer.from_config(ck.get("tokenizer_config", {}))
        encoder_params = dict(model_config.get("encoder", {}).get("cnn" if model_type == "cnn_lstm" else "resnet", {}))
        encoder_params["embedding_dim"] = model_config.get("embedding_dim", 256)
        model = Seq2SeqModel(model_type=model_type, vocab_size=tok.vocab_size, encoder_params=encoder_params,
                             decoder_params=model_config.get("decoder", {}), precision=precision)
        model.load_state_dict(ck["model_state_dict"])
        return cls(model, tok, device)

    def _prepare_image(self, image) -> torch.Tensor:
        """predictor.py:396-462 for one input; see `_prepare_images` for the batched form."""
        return self._prepare_images([image])

    def _prepare_images(self, images: Sequence) -> torch.Tensor:
        """predictor.py:396-462 applied to a chunk: (n, C, H, W) fp32 on the device.  Inputs of one kind are
        prepared together (one H2D copy + two launches for all paths / all PIL images of the chunk)."""
        enc = self.model.encoder
        size, channels = (enc.img_height, enc.img_width), enc.channels
        out: List[Optional[torch.Tensor]] = [None] * len(images)
        paths = [i for i, im in enumerate(images) if isinstance(im, str)]
        pils = [i for i, im in enumerate(images)
                if not isinstance(im, (str, torch.Tensor, np.ndarray)) and hasattr(im, "mode") and hasattr(im, "size")]
        if paths:                                                  # predictor.py:417-419 -> data/utils.py:18-90
            arrs = [P.open_image(images[i], channels) for i in paths]
            for mode_gray in (True, False):                        # a chunk may mix L and RGB files
                sel = [j for j, a in enumerate(arrs) if (a.ndim == 2) == mode_gray]
                if sel:
                    x = P.load_images([arrs[j] for j in sel], size, channels, device=self.device)
                    for j, row in zip(sel, x):
                        out[paths[j]] = row
        if pils:                                                   # predictor.py:427-446
            arrs = []
            for i in pils:
                im = images[i]
                if channels == 1 and im.mode != "L":
                    im = im.convert("L")
                elif channels == 3 and im.mode != "RGB":
                    im = im.convert("RGB")
                arrs.append(np.asarray(im))
            u8 = P.ResizePlan(arrs, size[0], size[1], resample="bicubic", mode="stretch").run(self.device)
            x = normalize_u8(u8, "pm1")
            for i, row in zip(pils, x):
                out[i] = row
        for i, im in enumerate(images):
            if out[i] is not None:
                continue
            if isinstance(im, np.ndarray):                         # predictor.py:423-426, 500-520
                a = im
                if a.ndim == 2:
                    a = a[None]
                elif a.ndim == 3 and a.shape[0] not in (1, 3):
                    a = np.transpose(a, (2, 0, 1))
                im = torch.from_numpy(np.ascontiguousarray(a)).float()
            if not isinstance(im, torch.Tensor):
                raise TypeError(f"Unsupported image type: {type(im)}. Expected str, torch.Tensor, numpy.ndarray, "
                                f"or PIL.Image.Image.")            # predictor.py:448-452
            x = im.to(self.device, dtype=torch.float32)
            if x.dim() == 2:
                x = x.unsqueeze(0)
            if x.dim() == 3:
                x = x.unsqueeze(0)
            if x.shape[-2:] != size:                               # predictor.py:483-492
                x = F.interpolate(x, size=size, mode="bilinear", align_corners=False)
            needs = (x.amin() < 0) | (x.amax() > 1)                 # predictor.py:494-497, decided on the device
            x = torch.where(needs, (x / 255.0) * 2.0 - 1.0, x)
            out[i] = x[0]
        rows = []
        for r in out:
            if self.model.model_type == "resnet_lstm" and r.shape[0] == 1:          # predictor.py:454-456
                r = r.repeat(3, 1, 1)
            rows.append(r)
        return torch.stack(rows)

    @torch.no_grad()
    def predict_batch(self, images: Union[torch.Tensor, Sequence[torch.Tensor]], beam_size: int = 0,
                      max_length: int = 141, temperature: float = 1.0, top_k: int = 0, top_p: float = 0.0,
                      batch_size: int = 16, seed: int = 0, uniforms: Optional[torch.Tensor] = None,
                      return_ids: bool = False) -> List[str]:
        if beam_size > 0:                                          # predictor.py:231-235
            logger.warning("Beam search is unsupported; using greedy decoding (beam_size=0).")
            beam_size = 0
        start, end = self.tokenizer.start_token_id, self.tokenizer.end_token_id
        results = []
        n = len(images)
        for i in range(0, n, batch_size):                          # predictor.py:240-246
            chunk = images[i:i + batch_size]
            batch = self._prepare_images(list(chunk))
            enc = self.model.encoder(batch)
            u = None if uniforms is None else uniforms[:, i:i + batch.shape[0]]
            tokens, lengths, steps = self.model.decoder.sample(enc, start, end, max_length, temperature, top_k, top_p,
                                                               seed=seed, offset=i * max_length, uniforms=u)
            toks, lens = tokens.tolist(), lengths.tolist()         # single host read per batch
            if lens and lens[0] < 0:                               # decode_wide.cu: the loop aborted on the device
                raise RuntimeError("the persistent decode loop aborted on the device (a bounded flag wait expired)")
            for row, ln in zip(toks, lens):
                seq = row[:ln]                                     # cut at first END (predictor.py:350-358)
                if seq and seq[0] == start:                        # predictor.py:384-385
                    seq = seq[1:]
                if seq and seq[-1] == end:
                    seq = seq[:-1]
                results.append(seq if return_ids else self.tokenizer.decode(seq))
        return results

    @torch.no_grad()
    def evaluate_batch(self, images, target_ids: torch.Tensor, max_length: Optional[int] = None, batch_size: int = 1024):
        """One batch of the ``img2latex evaluate`` loop (cli.py:448-495) kept on the device end to end: prepare ->
        encoder -> greedy ``predict_batch`` loop -> special-token filtering -> BLEU-4 / Levenshtein counts.  Returns
        the same dict as ``calculate_metrics``; the only device -> host read is the (B,8) count matrix."""
        from . import metrics as M
        max_length = self.tokenizer.max_sequence_length if max_length is None else max_length      # cli.py:458
        start, end = self.tokenizer.start_token_id, self.tokenizer.end_token_id
        counts = []
        for i in range(0, len(images), batch_size):
            batch = self._prepare_images(list(images[i:i + batch_size]))
            enc = self.model.encoder(batch)
            tokens, lengths, _ = self.model.decoder.sample(enc, start, end, max_length, 1.0, 0, 0.0)
            counts.append(M.evaluate_ids(tokens, lengths, target_ids[i:i + batch_size], self.tokenizer, return_counts=True))
        rows = torch.cat(counts).tolist()
        both = [M.scores_from_counts(r, 4) for r in rows]
        return {"bleu": sum(b for _, b in both) / len(rows), "levenshtein": sum(l for l, _ in both) / len(rows),
                "batch_size": len(rows)}

    @torch.no_grad()
    def predict(self, image: torch.Tensor, beam_size: int = 0, max_length: int = 141, temperature: float = 1.0,
                top_k: int = 0, top_p: float = 0.0) -> str:
        """predictor.py:139-203 (greedy through Seq2SeqModel.inference; beam clamped to 0)."""
        if beam_size > 0:
            logger.warning("Beam search is unsupported; using greedy decoding (beam_size=0).")
            beam_size = 0
        x = self._prepare_image(image)
        seq = self.model.inference(x, self.tokenizer.start_token_id, self.tokenizer.end_token_id, max_length,
                                   temperature, top_k, top_p, beam_size)
        if seq and seq[0] == self.tokenizer.start_token_id:        # predictor.py:194-198
            seq = seq[1:]
        if seq and seq[-1] == self.tokenizer.end_token_id:
            seq = seq[:-1]
        return self.tokenizer.decode(seq)
