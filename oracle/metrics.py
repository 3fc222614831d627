"""CPU oracle for the sequence metrics of `img2latex evaluate` (reference img2latex/training/metrics.py:
`levenshtein_distance` 49-94, `bleu_n_score` 97-179, `calculate_metrics` 182-223; called from cli.py:493-495).

TEST INFRASTRUCTURE ONLY (see oracle/__init__.py).  Pure-Python loops: use on small cases.  The integer parts
(`edit_distance`, `clipped_matches`) are what the CUDA kernel computes; `similarity_from_distance` /
`bleu_from_matches` are the float formulas the host applies to those integers.  Pinned against the live
reference by tests/golden/make_golden.py::metrics_case.
"""
from __future__ import annotations

import math
from typing import Dict, List, Sequence

__all__ = ["validation_loss_accuracy", "edit_distance", "clipped_matches", "similarity_from_distance", "bleu_from_matches",
           "levenshtein_distance", "bleu_n_score", "calculate_metrics"]


def edit_distance(a: Sequence[int], b: Sequence[int]) -> int:
    """metrics.py:60-85: the (rows+1) x (cols+1) table, unit costs, kept one row at a time."""
    prev = list(range(len(b) + 1))
    for r in range(1, len(a) + 1):
        cur = [r] + [0] * len(b)
        for c in range(1, len(b) + 1):
            if a[r - 1] == b[c - 1]:
                cur[c] = prev[c - 1]
            else:
                cur[c] = 1 + min(prev[c], cur[c - 1], prev[c - 1])
        prev = cur
    return prev[len(b)]


def similarity_from_distance(dist: int, rows: int, cols: int) -> float:
    """metrics.py:87-94."""
    max_length = max(rows, cols)
    if max_length == 0:
        return 1.0
    return 1.0 - (dist / max_length)


def levenshtein_distance(a: Sequence[int], b: Sequence[int]) -> float:
    return similarity_from_distance(edit_distance(a, b), len(a), len(b))


def clipped_matches(gen: Sequence[int], true: Sequence[int], gram_size: int) -> int:
    """metrics.py:131-152: sum over the distinct n-grams of the generated sequence of min(count_gen, count_true)."""
    if len(gen) < gram_size or len(true) < gram_size:
        return 0
    cg: Dict[tuple, int] = {}
    ct: Dict[tuple, int] = {}
    for i in range(len(gen) - gram_size + 1):
        g = tuple(gen[i:i + gram_size])
        cg[g] = cg.get(g, 0) + 1
    for i in range(len(true) - gram_size + 1):
        g = tuple(true[i:i + gram_size])
        ct[g] = ct.get(g, 0) + 1
    return sum(min(v, ct.get(g, 0)) for g, v in cg.items())


def bleu_from_matches(matches: Sequence[int], gen_len: int, true_len: int, n: int) -> float:
    """metrics.py:114-179 given the clipped match counts for gram sizes 1..n."""
    if gen_len == 0 or true_len == 0:
        return 0.0
    scores = []
    for gram_size in range(1, n + 1):
        if gen_len < gram_size or true_len < gram_size:
            scores.append(0.0)
            continue
        scores.append(matches[gram_size - 1] / (gen_len - gram_size + 1))
    for s in scores:
        if s == 0.0:
            return 0.0
    geo_mean = 0.0
    for s in scores:
        geo_mean += math.log(s)
    geo_mean = math.exp(geo_mean / n)
    if gen_len < true_len:
        return math.exp(1.0 - true_len / gen_len) * geo_mean
    return geo_mean


def bleu_n_score(gen: Sequence[int], true: Sequence[int], n: int = None) -> float:
    n = 4 if n is None else n
    return bleu_from_matches([clipped_matches(gen, true, g) for g in range(1, n + 1)], len(gen), len(true), n)


def calculate_metrics(predictions: List[List[int]], targets: List[List[int]]) -> Dict[str, float]:
    """metrics.py:182-223."""
    assert len(predictions) == len(targets)
    num = len(predictions)
    bleu = [bleu_n_score(predictions[i], targets[i], 4) for i in range(num)]
    lev = [levenshtein_distance(predictions[i], targets[i]) for i in range(num)]
    return {"bleu": sum(bleu) / num, "levenshtein": sum(lev) / num, "batch_size": num}


def validation_loss_accuracy(logits, targets, pad_token_id: int, label_smoothing: float = 0.1):
    """training/trainer.py:111-115 + 517-529 and training/metrics.py:226-238 restated with torch functional ops on the
    CPU (test infrastructure): loss = CrossEntropyLoss(ignore_index=pad, reduction="mean", label_smoothing) over
    logits (B,T,V) / targets (B,T); (correct, total) = masked_accuracy."""
    import torch
    import torch.nn.functional as F
    loss = F.cross_entropy(logits.transpose(1, 2), targets, ignore_index=pad_token_id, reduction="mean",
                           label_smoothing=label_smoothing)
    pred = torch.argmax(logits, dim=-1)
    mask = targets.ne(pad_token_id)
    return loss, int(torch.logical_and(pred.eq(targets), mask).sum()), int(mask.sum())
