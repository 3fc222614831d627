"""Batch-sharded data-parallel inference: images are independent through
encoder and decode, so each rank takes a contiguous slice of the batch with a full weight
replica and NO per-step communication.  The one exchange is an all-gather of the
(B/G, T+1) token ids (+ lengths) at the end; the global stop step of the sticky rule
composes as max over ranks.  Works with NCCL (GPU) and gloo (CPU tests)."""
from __future__ import annotations

from typing import List, Tuple

import torch
import torch.distributed as dist


def shard_bounds(n: int, world_size: int, rank: int) -> Tuple[int, int]:
    """Contiguous, balanced slices; the first n % world_size ranks get one extra item."""
    base, rem = divmod(n, world_size)
    lo = rank * base + min(rank, rem)
    return lo, lo + base + (1 if rank < rem else 0)


def gather_tokens(tokens: torch.Tensor, lengths: torch.Tensor, steps: torch.Tensor, n_total: int,
                  pad_id: int = -1) -> Tuple[torch.Tensor, torch.Tensor, torch.Tensor]:
    """All-gather ragged per-rank (b_r, T+1) token matrices into the (n_total, T+1) global
    matrix in rank order.  Every rank receives the full result."""
    if not dist.is_initialized() or dist.get_world_size() == 1:
        return tokens, lengths, steps
    ws = dist.get_world_size()
    T1 = tokens.shape[1]
    cap = (n_total + ws - 1) // ws
    # ONE collective: rows [0, cap) = (tokens | length) of the shard, row cap = the rank's stop step
    b = tokens.shape[0]
    buf = torch.empty((cap + 1, T1 + 1), dtype=torch.int64, device=tokens.device)
    if b < cap:
        buf[b:cap].fill_(pad_id)
    buf[:b, :T1] = tokens
    buf[:b, T1] = lengths
    buf[cap] = steps
    out = torch.empty(ws, cap + 1, T1 + 1, dtype=torch.int64, device=tokens.device)
    dist.all_gather_into_tensor(out.view(ws * (cap + 1), T1 + 1), buf)
    gsteps = out[:, cap, 0].max().to(torch.int32)                  # the sticky stop composes as max over ranks
    if n_total == ws * cap:
        full = out[:, :cap].reshape(n_total, T1 + 1)
    else:
        rows: List[torch.Tensor] = []
        for r in range(ws):
            lo, hi = shard_bounds(n_total, ws, r)
            rows.append(out[r, : hi - lo])
        full = torch.cat(rows, dim=0)
    return full[:, :T1], full[:, T1].to(torch.int32), gsteps        # tokens: a view of the gathered buffer


def gather_counts(counts: torch.Tensor, n_total: int) -> torch.Tensor:
    """Sharded `img2latex evaluate`: all-gather the per-rank (b_r, 8) int32 count matrices of
    `metrics.sequence_counts` / `metrics.evaluate_ids` into the (n_total, 8) matrix in global image order, so that
    `metrics.scores_from_counts` + the running sums of `calculate_metrics` (training/metrics.py:205-218) see the pairs
    in the same order as a single process does (the means are then bit-identical)."""
    if not dist.is_initialized() or dist.get_world_size() == 1:
        return counts
    ws = dist.get_world_size()
    cap = (n_total + ws - 1) // ws
    b = counts.shape[0]
    buf = torch.zeros((cap, counts.shape[1]), dtype=counts.dtype, device=counts.device)
    buf[:b] = counts
    out = torch.empty(ws * cap, counts.shape[1], dtype=counts.dtype, device=counts.device)
    dist.all_gather_into_tensor(out, buf)
    rows = []
    for r in range(ws):
        lo, hi = shard_bounds(n_total, ws, r)
        rows.append(out[r * cap: r * cap + (hi - lo)])
    return torch.cat(rows, dim=0)


def reduce_validation(loss: torch.Tensor, correct: int, tokens: int) -> Tuple[float, int, int]:
    """Sharded validation step: every rank holds the token-mean loss of its shard (`metrics.cross_entropy_metrics`,
    reduction "mean" over non-pad tokens) and its (correct, tokens) counts.  The global mean over tokens is the
    token-weighted mean of the shard means (both terms of the label-smoothed loss are normalised by the same
    count); combined in fp64.  Returns (loss, correct, tokens) of the whole batch on every rank."""
    if not dist.is_initialized() or dist.get_world_size() == 1:
        return float(loss), int(correct), int(tokens)
    v = torch.tensor([float(loss) * tokens if tokens else 0.0, float(correct), float(tokens)], dtype=torch.float64,
                     device=loss.device)
    dist.all_reduce(v, op=dist.ReduceOp.SUM)
    tot = int(v[2].item())
    return (float(v[0].item()) / tot if tot else float("nan")), int(v[1].item()), tot
