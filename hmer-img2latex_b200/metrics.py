"""Evaluation metrics of ``img2latex evaluate`` (reference img2latex/training/metrics.py: ``levenshtein_distance``
49-94, ``bleu_n_score`` 97-179, ``calculate_metrics`` 182-223; called from cli.py:493-495) with the
O(T^2) integer work -- the edit-distance table and the clipped n-gram matching -- on the device
(``i2l_sequence_metrics``, one warp per (prediction, target) pair).  The float formulas are applied on the
host to the returned integers with the same ``math`` calls as the reference, so every score is bit-identical
to the reference's Python result.  No CPU fallback: without the CUDA library every call raises."""
from __future__ import annotations

import ctypes as C
import math
from typing import Dict, List, Optional, Sequence, Tuple

import torch

from . import _native as N

__all__ = ["cross_entropy_metrics", "masked_accuracy", "filter_ids", "evaluate_ids", "sequence_counts", "levenshtein_distance", "bleu_n_score", "calculate_metrics", "scores_from_counts"]


def _pad(seqs: Sequence[Sequence[int]]) -> Tuple[torch.Tensor, torch.Tensor]:
    n = max((len(s) for s in seqs), default=0)
    ids = torch.zeros(len(seqs), max(n, 1), dtype=torch.int64)
    for i, s in enumerate(seqs):
        if len(s):
            ids[i, : len(s)] = torch.as_tensor(list(s), dtype=torch.int64)
    return ids, torch.tensor([len(s) for s in seqs], dtype=torch.int32)


def sequence_counts(pred: torch.Tensor, pred_len: torch.Tensor, tgt: torch.Tensor, tgt_len: torch.Tensor,
                    max_n: int = 4) -> torch.Tensor:
    """Device entry: pred (B,Tp) / tgt (B,Tt) int64 id matrices with their lengths (e.g. the token matrix a decode
    loop just produced) -> int32 (B,8): [edit distance, matches n=1..4, pred_len, tgt_len, 0].  Stream-ordered,
    no host sync."""
    if not (pred.is_cuda and tgt.is_cuda):
        raise RuntimeError("sequence_counts needs CUDA tensors; there is no CPU fallback")
    dev = pred.device
    pred, tgt = pred.to(torch.int64).contiguous(), tgt.to(device=dev, dtype=torch.int64).contiguous()
    pred_len = pred_len.to(device=dev, dtype=torch.int32).contiguous()
    tgt_len = tgt_len.to(device=dev, dtype=torch.int32).contiguous()
    B = pred.shape[0]
    if tgt.shape[0] != B or pred_len.numel() != B or tgt_len.numel() != B:
        raise AssertionError("Predictions and targets must have the same length")    # metrics.py:200-202
    out = torch.zeros(B, 8, dtype=torch.int32, device=dev)
    with torch.cuda.device(dev):
        N.check(N.lib().i2l_sequence_metrics(N.ptr(pred), pred.shape[1], N.ptr(pred_len), N.ptr(tgt), tgt.shape[1],
                                             N.ptr(tgt_len), B, int(max_n), N.ptr(out), N.stream_ptr(dev)),
                "i2l_sequence_metrics")
    return out


def cross_entropy_metrics(logits: torch.Tensor, targets: torch.Tensor, pad_token_id: int, label_smoothing: float = 0.1,
                          sync: bool = True):
    """Validation step of the reference trainer (training/trainer.py:517-529) on device-resident logits:
    ``nn.CrossEntropyLoss(ignore_index=pad, reduction="mean", label_smoothing=0.1)`` (trainer.py:111-115) and
    ``masked_accuracy`` (training/metrics.py:226-238) in one pass over the logits.  logits (B,T,V) fp32 (what
    ``Seq2SeqModel.forward`` returns), targets (B,T) int64.  Returns (loss tensor, correct, tokens) -- with
    ``sync=False`` correct / tokens stay on the device as an int32[4] tensor."""
    if not logits.is_cuda:
        raise RuntimeError("cross_entropy_metrics needs CUDA tensors; there is no CPU fallback")
    dev = logits.device
    V = logits.shape[-1]
    lg = logits.to(torch.float32).contiguous().view(-1, V)
    tg = targets.to(device=dev, dtype=torch.int64).contiguous().view(-1)
    if tg.numel() != lg.shape[0]:
        raise ValueError(f"Expected {lg.shape[0]} targets, got {tg.numel()}")
    n = lg.shape[0]
    loss = torch.empty((), dtype=torch.float32, device=dev)
    counts = torch.zeros(4, dtype=torch.int32, device=dev)
    lib = N.lib()
    ws = torch.empty(max(lib.i2l_xent_workspace_bytes(n), 256), dtype=torch.uint8, device=dev)
    with torch.cuda.device(dev):
        N.check(lib.i2l_xent_metrics(N.ptr(lg), N.ptr(tg), n, V, int(pad_token_id), float(label_smoothing), N.ptr(loss),
                                     N.ptr(counts), N.ptr(ws), ws.numel(), N.stream_ptr(dev)), "i2l_xent_metrics")
    if not sync:
        return loss, counts
    c = counts.tolist()
    if c[2]:
        raise IndexError("Target out of bounds")            # what torch's cross_entropy raises on the CPU
    return loss, c[0], c[1]


def masked_accuracy(logits: torch.Tensor, targets: torch.Tensor, pad_token_id: int) -> Tuple[int, int]:
    """training/metrics.py:226-238: (correct, total) over the non-pad positions."""
    _, correct, total = cross_entropy_metrics(logits, targets, pad_token_id, 0.0)
    return correct, total


def filter_ids(ids: torch.Tensor, lengths: Optional[torch.Tensor], drop: Sequence[int]) -> Tuple[torch.Tensor, torch.Tensor]:
    """Per-row stable compaction on the device: rows of ``ids[:, :len]`` without the ids in ``drop`` (<= 8 of them)
    -> (filtered (B,T) int64, new lengths (B) int32)."""
    if not ids.is_cuda:
        raise RuntimeError("filter_ids needs CUDA tensors; there is no CPU fallback")
    dev = ids.device
    ids = ids.to(torch.int64).contiguous()
    B, T = ids.shape
    out = torch.zeros(B, max(T, 1), dtype=torch.int64, device=dev)
    out_len = torch.zeros(B, dtype=torch.int32, device=dev)
    ln = None if lengths is None else lengths.to(device=dev, dtype=torch.int32).contiguous()
    d = (C.c_int64 * max(len(drop), 1))(*[int(x) for x in drop])
    with torch.cuda.device(dev):
        N.check(N.lib().i2l_filter_ids(N.ptr(ids), T, N.ptr(ln), B, d, len(drop), N.ptr(out), out.shape[1], N.ptr(out_len),
                                       N.stream_ptr(dev)), "i2l_filter_ids")
    return out, out_len


def evaluate_ids(pred_tokens: torch.Tensor, pred_len: Optional[torch.Tensor], target_ids: torch.Tensor, tokenizer,
                 return_counts: bool = False):
    """The metric part of ``img2latex evaluate`` (cli.py:448-495) on device-resident ids, without the detour through
    strings: predictions lose PAD/START/END/UNK (what ``decode`` + ``encode`` do to them, cli.py:476-479), targets
    lose PAD only (cli.py:471-474); then mean BLEU-4 / Levenshtein similarity as ``calculate_metrics``."""
    specials = [tokenizer.pad_token_id, tokenizer.start_token_id, tokenizer.end_token_id, tokenizer.unk_token_id]
    p, pl = filter_ids(pred_tokens, pred_len, specials)
    t, tl = filter_ids(target_ids.to(pred_tokens.device), None, [tokenizer.pad_token_id])
    counts = sequence_counts(p, pl, t, tl, 4)
    if return_counts:
        return counts
    rows = counts.tolist()                                 # the one device -> host read
    both = [scores_from_counts(r, 4) for r in rows]
    num = len(rows)
    return {"bleu": sum(b for _, b in both) / num, "levenshtein": sum(l for l, _ in both) / num, "batch_size": num}


def scores_from_counts(row: Sequence[int], n: int = 4) -> Tuple[float, float]:
    """(levenshtein similarity, BLEU-n) of one pair from its integer counts: metrics.py:87-94 and 114-179."""
    dist, matches, gen_len, true_len = row[0], row[1:5], row[5], row[6]
    max_length = max(gen_len, true_len)
    lev = 1.0 if max_length == 0 else 1.0 - (dist / max_length)
    if gen_len == 0 or true_len == 0:
        return lev, 0.0
    scores = []
    for gram_size in range(1, n + 1):
        if gen_len < gram_size or true_len < gram_size:
            scores.append(0.0)
        else:
            scores.append(matches[gram_size - 1] / (gen_len - gram_size + 1))
    if any(s == 0.0 for s in scores):
        return lev, 0.0
    geo_mean = 0.0
    for s in scores:
        geo_mean += math.log(s)
    geo_mean = math.exp(geo_mean / n)
    if gen_len < true_len:
        return lev, math.exp(1.0 - true_len / gen_len) * geo_mean
    return lev, geo_mean


def _counts_host(predictions: Sequence[Sequence[int]], targets: Sequence[Sequence[int]], max_n: int, device) -> List[List[int]]:
    device = torch.device(device) if device is not None else torch.device("cuda", torch.cuda.current_device())
    p, pl = _pad(predictions)
    t, tl = _pad(targets)
    out = sequence_counts(p.to(device), pl.to(device), t.to(device), tl.to(device), max_n)
    return out.tolist()                                   # the one device -> host read


def levenshtein_distance(sequence_one: Sequence[int], sequence_two: Sequence[int], device=None) -> float:
    """metrics.py:49-94 (returns the normalised similarity, like the reference)."""
    return scores_from_counts(_counts_host([sequence_one], [sequence_two], 1, device)[0], 1)[0]


def bleu_n_score(generated_sequence: Sequence[int], true_sequence: Sequence[int], n: Optional[int] = None, device=None) -> float:
    """metrics.py:97-179."""
    n = 4 if n is None else n
    if not 1 <= n <= 4:
        raise ValueError("n-gram sizes 1..4 are supported on the device path")
    return scores_from_counts(_counts_host([generated_sequence], [true_sequence], n, device)[0], n)[1]


def calculate_metrics(predictions: Sequence[Sequence[int]], targets: Sequence[Sequence[int]], device=None) -> Dict[str, float]:
    """metrics.py:182-223: mean BLEU-4 and mean Levenshtein similarity of a set of sequences -- one launch for
    the whole set."""
    predictions = [p.tolist() if isinstance(p, torch.Tensor) else list(p) for p in predictions]
    targets = [t.tolist() if isinstance(t, torch.Tensor) else list(t) for t in targets]
    assert len(predictions) == len(targets), "Predictions and targets must have the same length"
    num_sequences = len(predictions)
    rows = _counts_host(predictions, targets, 4, device)
    both = [scores_from_counts(r, 4) for r in rows]
    mean_bleu = sum(b for _, b in both) / num_sequences          # ZeroDivisionError on an empty set, like the reference
    mean_lev = sum(l for l, _ in both) / num_sequences
    return {"bleu": mean_bleu, "levenshtein": mean_lev, "batch_size": num_sequences}
