"""Batch-sharded data-parallel inference: images are independent through
encoder and decode, so each rank takes a contiguous slice of the batch with a full weight
replica and NO per-step communication.  The one exchange is an all-gather of the
(B/G, T+1) token ids (+ lengths) at the end; the global stop step of the sticky rule
composes as max over ranks.  Works with NCCL (GPU) and gloo (CPU tests)."""
from __future__ import annotations

import ctypes as C
from typing import List, Optional, Tuple

import torch
import torch.distributed as dist

from . import _native as N


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


class TokenExchange:
    """The token all-gather as direct peer stores (csrc/token_exchange.cu, `i2l_token_exchange_*`): every rank owns a
    receive buffer in symmetric memory (`torch.distributed._symmetric_memory`: cudaMalloc'd, peer-mapped over NVLink);
    `write()` stores the rank's shard into its slot of EVERY peer's buffer and publishes a sequence flag, `read()`
    acquires the flags of the local buffer and returns the global (n_total, T1) matrix -- no library collective, no
    host synchronisation, nothing on the compute stream but two small kernels per step.

    Pipelined use (what `bench.py --gpus N` and a serving loop do): per step `prev = xchg.step(tokens, lengths, steps)`
    = read(previous step) then write(this step); the result of step i is returned by the call of step i + 1 (or by
    `flush()`), a whole step after the peers produced it, so a slow rank does not stall the others' compute.
    `peers=None` builds a single-process exchange over explicitly given buffers (tests: several "ranks" on one GPU).
    `gather_tokens` (NCCL / gloo `all_gather_into_tensor`) stays as the checked reference of this path."""

    # How far the compute stream may run ahead of the exchange stream (kicks).  Measured at N = 8, 20 timed steps of
    # 0.92 ms: 3 -> 0.997 ms per step, 32 -> 1.067 ms (the exchange kernels of many steps then pile up behind the
    # persistent compute kernels and spill into the encoders, whose one-CTA-per-SM tile ranges they delay).
    RUN_AHEAD = 3
    fence_after_decode = True

    def __init__(self, n_total: int, T1: int, device: torch.device, group=None, rank: Optional[int] = None,
                 world: Optional[int] = None, buffers: Optional[List[torch.Tensor]] = None):
        self.n_total, self.T1, self.device = int(n_total), int(T1), torch.device(device)
        lib = N.lib()
        if buffers is None:
            import torch.distributed._symmetric_memory as symm_mem
            group = group or dist.group.WORLD
            self.rank, self.world = dist.get_rank(group), dist.get_world_size(group)
            nbytes = lib.i2l_token_exchange_buffer_bytes(self.world, self.n_total, self.T1)
            self.local = symm_mem.empty(nbytes, dtype=torch.uint8, device=self.device)
            self.local.zero_()
            torch.cuda.synchronize(self.device)
            self._hdl = symm_mem.rendezvous(self.local, group)
            ptrs = [int(p) for p in self._hdl.buffer_ptrs]
            self._hdl.barrier()                                   # every buffer is zeroed before anyone writes
        else:
            self.rank, self.world = int(rank), int(world)
            nbytes = lib.i2l_token_exchange_buffer_bytes(self.world, self.n_total, self.T1)
            assert len(buffers) == self.world and all(b.numel() >= nbytes for b in buffers)
            self.local, self._hdl = buffers[self.rank], None
            ptrs = [b.data_ptr() for b in buffers]
        if nbytes == 0:
            raise RuntimeError("i2l_token_exchange_buffer_bytes rejected the configuration")
        self._ptrs = (C.c_void_p * self.world)(*ptrs)
        self.seq = 0                                              # sequence number of the last write
        self.read_seq = 0
        self._timeout = torch.zeros((), dtype=torch.int32, device=self.device)
        self._stream = None                                       # side stream of kick() / flush(), created on first use
        self._staged = None
        self.result_event = None
        lo, hi = shard_bounds(self.n_total, self.world, self.rank)
        self.shard, self.shard_lo = hi - lo, lo

    def write(self, tokens: torch.Tensor, lengths: torch.Tensor, steps: torch.Tensor) -> None:
        if tuple(tokens.shape) != (self.shard, self.T1) or tokens.dtype != torch.int64 or not tokens.is_contiguous():
            raise RuntimeError(f"TokenExchange.write expects contiguous int64 ({self.shard},{self.T1}) tokens, got "
                               f"{tokens.dtype} {tuple(tokens.shape)}")
        if self.seq != self.read_seq:
            raise RuntimeError("TokenExchange: read() the previous step before writing the next (two parities)")
        self.seq += 1
        with torch.cuda.device(self.device):
            N.check(N.lib().i2l_token_exchange_write(N.ptr(tokens), N.ptr(lengths.to(torch.int32)),
                                                     N.ptr(steps.to(torch.int32)), self.shard, self.T1, self.n_total,
                                                     self.rank, self.world, self._ptrs, self.seq,
                                                     N.stream_ptr(self.device)), "i2l_token_exchange_write")

    def read(self, out: Optional[Tuple[torch.Tensor, torch.Tensor, torch.Tensor]] = None
             ) -> Tuple[torch.Tensor, torch.Tensor, torch.Tensor]:
        """Global (tokens (n_total,T1) int64, lengths (n_total) int32, steps) of the last written step, on the current
        stream; `out` = caller-owned result tensors (otherwise fresh ones)."""
        if self.read_seq == self.seq:
            raise RuntimeError("TokenExchange.read: nothing pending")
        self.read_seq = self.seq
        if out is None:
            out = (torch.empty(self.n_total, self.T1, dtype=torch.int64, device=self.device),
                   torch.empty(self.n_total, dtype=torch.int32, device=self.device),
                   torch.empty((), dtype=torch.int32, device=self.device))
        tokens, lengths, steps = out
        with torch.cuda.device(self.device):
            N.check(N.lib().i2l_token_exchange_read(N.ptr(self.local), self.world, self.n_total, self.T1, self.seq,
                                                    N.ptr(tokens), N.ptr(lengths), N.ptr(steps), N.ptr(self._timeout),
                                                    N.stream_ptr(self.device)), "i2l_token_exchange_read")
        return tokens, lengths, steps

    # ---- pipelined use: the exchange runs on its OWN stream, next to the decode kernel of the following batch ----------
    # The encoder kernels are persistent with one CTA per SM, so anything that shares the GPU with them delays whole
    # tile ranges (measured: conv2 0.16 -> 0.27 ms with the exchange beside it); the persistent decode kernel leaves 20
    # of the 148 SMs idle and is latency-bound.  `stage()` therefore only remembers a step's result; `kick()`, called
    # by the serving loop between the NEXT batch's encoder and decode, releases it on the side stream.
    def _side(self):
        if self._stream is None:
            self._stream = torch.cuda.Stream(self.device)
            self._produced = [torch.cuda.Event(), torch.cuda.Event()]
            self._consumed = [torch.cuda.Event() for _ in range(self.RUN_AHEAD)]
            self._here = [torch.cuda.Event() for _ in range(4)]
            self._held = [None] * (self.RUN_AHEAD + 1)
            self.result_event = torch.cuda.Event()
            # three rotating result sets: a result handed out by kick() stays valid for two more kicks
            self._ring = [(torch.empty(self.n_total, self.T1, dtype=torch.int64, device=self.device),
                           torch.empty(self.n_total, dtype=torch.int32, device=self.device),
                           torch.empty((), dtype=torch.int32, device=self.device)) for _ in range(3)]
            self._turn = 0
            self._kicks = 0
        return self._stream

    def stage(self, tokens, lengths, steps) -> None:
        """remember this step's shard result (produced on the current stream) for the next kick()"""
        if self._staged is not None:
            raise RuntimeError("TokenExchange.stage: kick() the previous step first")
        side = self._side()
        cur = torch.cuda.current_stream(self.device)
        ev = self._produced[self.seq & 1]
        ev.record(cur)
        if self.fence_after_decode and self._kicks > 0:
            # The exchange released by the last kick() runs beside the decode kernel that has just been enqueued; whatever
            # the current stream runs NEXT (the following batch's encoder: persistent kernels with one CTA per SM) must not
            # share the GPU with it.  Normally the ~60 us of exchange are long over when the 390 us decode ends and this
            # wait costs nothing; when a peer is late the encoder waits instead of being slowed down for its whole run --
            # measured at N = 8: exchange kernels that slip under the encoders put every rank into a state where each
            # step's exchange is late for the next one (1.42 ms per step instead of 1.00).
            cur.wait_event(self._consumed[(self._kicks - 1) % self.RUN_AHEAD])
        self._staged = (tokens, lengths, steps, ev)

    def kick(self):
        """read(previous step) then write(staged step) on the side stream, ordered after the work queued on the current
        stream so far (the serving loop calls it right before the decode launch of the next batch).  Returns the
        previous step's global result or None: device tensors of a 3-deep ring, complete once `result_event` has fired
        -- `wait_result()` / `flush()` make the current stream wait for it.  The current stream is held back only when
        it runs more than RUN_AHEAD exchanges ahead of the side stream (a peer that far behind)."""
        if self._staged is None:
            return None
        side = self._side()
        cur = torch.cuda.current_stream(self.device)
        tokens, lengths, steps, produced = self._staged
        self._staged = None
        here = self._here[self._kicks % 4]
        here.record(cur)
        if self._kicks >= self.RUN_AHEAD:                      # bounded run-ahead: the kick of RUN_AHEAD steps ago is done
            cur.wait_event(self._consumed[self._kicks % self.RUN_AHEAD])
        prev = None
        with torch.cuda.stream(side):
            side.wait_event(produced)
            side.wait_event(here)
            if self.read_seq != self.seq:
                prev = self.read(self._ring[self._turn])
                self._turn = (self._turn + 1) % 3
                self.result_event.record(side)
            self.write(tokens, lengths, steps)
            self._consumed[self._kicks % self.RUN_AHEAD].record(side)
        # The inputs stay referenced for RUN_AHEAD + 1 more kicks instead of record_stream(): by then the current stream has
        # waited for this kick's `_consumed` event, so whatever reuses their memory is ordered after the peer stores.
        # (record_stream defers the reuse to the caching allocator's event polling; its pool then grows by cudaMalloc --
        # slow and device-synchronising with peer mappings in place -- for the first dozens of steps of every run.)
        self._held[self._kicks % (self.RUN_AHEAD + 1)] = (tokens, lengths, steps)
        self._kicks += 1
        return prev

    def step(self, tokens, lengths, steps):
        """stage() + kick() at once: the exchange starts as soon as the step's result exists (beside whatever the
        current stream runs next).  Returns the previous step's global result (None at first), see kick()."""
        self.stage(tokens, lengths, steps)
        return self.kick()

    def wait_result(self) -> None:
        """make the current stream wait for the last result handed out by kick() / step()"""
        if self._stream is not None:
            torch.cuda.current_stream(self.device).wait_event(self.result_event)

    def flush(self):
        """Drain the pipeline: the results not handed out yet, oldest first (0, 1 or 2 triples); the current stream
        waits for them."""
        out = []
        if self._staged is not None:
            prev = self.kick()
            if prev is not None:
                out.append(prev)
        if self.read_seq != self.seq:
            if self._stream is None:
                out.append(self.read())
            else:
                side = self._side()
                with torch.cuda.stream(side):
                    out.append(self.read(self._ring[self._turn]))
                    self._turn = (self._turn + 1) % 3
                    self.result_event.record(side)
        self.wait_result()
        return out

    def check(self) -> None:
        """Host-side check (synchronises): raises if a read kernel gave up waiting for a peer."""
        if int(self._timeout.item()):
            raise RuntimeError("TokenExchange: a peer's shard did not arrive within the spin bound")


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
