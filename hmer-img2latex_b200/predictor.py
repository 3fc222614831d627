"""Serving facade mirroring ``img2latex/training/predictor.py`` (Predictor 28-521) for the
hot path: ``predict`` / ``predict_batch`` on image *tensors* with the reference's loop
semantics (sampling only when temperature > 0 and (top_k > 0 or top_p > 0), sticky
finished flags, cut at the first END, beam clamped to greedy -- predictor.py:163-167,
231-235).  File / PIL loading (``_prepare_image`` 396-462) is host-side I/O outside the
hot path; tensors of the wrong size are resized bilinearly as in ``_preprocess_tensor``
(464-499) but to the MODEL's size (the reference hard-codes 64x800)."""
from __future__ import annotations

import logging
from typing import List, Optional, Sequence, Union

import torch
import torch.nn.functional as F

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
    def from_checkpoint(cls, checkpoint_path: str, device=None) -> "Predictor":
        """Checkpoint layout of training/trainer.py:209-221, consumed as in predictor.py:61-137."""
        ck = torch.load(checkpoint_path, map_location="cpu", weights_only=False)
        mc = ck["config"]["model"]
        tok = LaTeXTokenizer.from_config(ck["tokenizer_config"])
        model = Seq2SeqModel(mc["name"], tok.vocab_size, dict(mc.get("encoder", {}), embedding_dim=mc.get("embedding_dim", 256)),
                             mc.get("decoder", {}))
        model.load_state_dict(ck["model_state_dict"])
        return cls(model, tok, device)

    def _prepare_image(self, image: torch.Tensor) -> torch.Tensor:
        if not isinstance(image, torch.Tensor):
            raise TypeError(f"Unsupported image type: {type(image)} (tensor inputs only on the hot path)")  # cf. predictor.py:448-452
        enc = self.model.encoder
        x = image.to(self.device, dtype=torch.float32)
        if x.dim() == 3:
            x = x.unsqueeze(0)
        if x.shape[-2:] != (enc.img_height, enc.img_width):
            x = F.interpolate(x, size=(enc.img_height, enc.img_width), mode="bilinear", align_corners=False)
        return x

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
            batch = torch.cat([self._prepare_image(im) for im in chunk], dim=0)
            enc = self.model.encoder(batch)
            u = None if uniforms is None else uniforms[:, i:i + batch.shape[0]]
            tokens, lengths, steps = self.model.decoder.sample(enc, start, end, max_length, temperature, top_k, top_p,
                                                               seed=seed, offset=i * max_length, uniforms=u)
            toks, lens = tokens.tolist(), lengths.tolist()         # single host read per batch
            for row, ln in zip(toks, lens):
                seq = row[:ln]                                     # cut at first END (predictor.py:350-358)
                if seq and seq[0] == start:                        # predictor.py:384-385
                    seq = seq[1:]
                if seq and seq[-1] == end:
                    seq = seq[:-1]
                results.append(seq if return_ids else self.tokenizer.decode(seq))
        return results

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
