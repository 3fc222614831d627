"""Token exchange kernels (csrc/token_exchange.cu) on ONE GPU: several "ranks" live in one process with one receive
buffer each, every rank writes its shard into all buffers and every rank's read must reproduce the global matrix --
the layout / flag / parity logic of the multi-GPU path (peer-mapped buffers over NVLink differ only in the address)."""
import pytest
import torch

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("world,n_total,T1", [(1, 5, 7), (2, 8, 151), (4, 10, 33), (8, 1024, 151), (3, 2, 5)])
def test_exchange_single_process_ranks(pkg, world, n_total, T1):
    from hmer_img2latex_b200.dist import TokenExchange, shard_bounds
    N = pkg._native
    dev = torch.device("cuda", 0)
    nbytes = N.lib().i2l_token_exchange_buffer_bytes(world, n_total, T1)
    bufs = [torch.zeros(nbytes, dtype=torch.uint8, device=dev) for _ in range(world)]
    xs = [TokenExchange(n_total, T1, dev, rank=r, world=world, buffers=bufs) for r in range(world)]
    g = torch.Generator().manual_seed(world * 100 + n_total)
    for it in range(5):                                    # both parities, several rounds of buffer reuse
        full = torch.randint(0, 512, (n_total, T1), generator=g, dtype=torch.int64)
        lens = torch.randint(1, T1 + 1, (n_total,), generator=g, dtype=torch.int32)
        steps = [int(torch.randint(1, 150, (1,), generator=g)) for _ in range(world)]
        for r, x in enumerate(xs):
            lo, hi = shard_bounds(n_total, world, r)
            x.write(full[lo:hi].contiguous().to(dev), lens[lo:hi].contiguous().to(dev),
                    torch.tensor(steps[r], dtype=torch.int32, device=dev))
        for x in xs:
            tok, ln, st = x.read()
            assert torch.equal(tok.cpu(), full) and torch.equal(ln.cpu(), lens) and int(st) == max(steps)
            x.check()


def test_exchange_read_times_out_instead_of_hanging(pkg):
    """A peer that never writes: the read kernel's bounded spin sets the timeout flag and returns."""
    from hmer_img2latex_b200.dist import TokenExchange
    N = pkg._native
    dev = torch.device("cuda", 0)
    nbytes = N.lib().i2l_token_exchange_buffer_bytes(2, 4, 3)
    bufs = [torch.zeros(nbytes, dtype=torch.uint8, device=dev) for _ in range(2)]
    x0 = TokenExchange(4, 3, dev, rank=0, world=2, buffers=bufs)
    x0.write(torch.zeros(2, 3, dtype=torch.int64, device=dev), torch.ones(2, dtype=torch.int32, device=dev),
             torch.tensor(1, dtype=torch.int32, device=dev))
    x0.read()
    with pytest.raises(RuntimeError, match="did not arrive"):
        x0.check()


@pytest.mark.parametrize("world,n_total,T1", [(2, 8, 151), (4, 1024, 151)])
def test_exchange_pipelined_on_side_streams(pkg, world, n_total, T1):
    """TokenExchange.step / flush: read(previous) + write(current) run on the exchange's own stream behind the compute
    stream; step i hands out the global result of step i - 1 (valid after wait_result), flush the last one."""
    from hmer_img2latex_b200.dist import TokenExchange, shard_bounds
    N = pkg._native
    dev = torch.device("cuda", 0)
    nbytes = N.lib().i2l_token_exchange_buffer_bytes(world, n_total, T1)
    bufs = [torch.zeros(nbytes, dtype=torch.uint8, device=dev) for _ in range(world)]
    xs = [TokenExchange(n_total, T1, dev, rank=r, world=world, buffers=bufs) for r in range(world)]
    g = torch.Generator().manual_seed(7 * world + n_total)
    history = []
    for it in range(6):
        full = torch.randint(0, 512, (n_total, T1), generator=g, dtype=torch.int64)
        lens = torch.randint(1, T1 + 1, (n_total,), generator=g, dtype=torch.int32)
        steps = [int(torch.randint(1, 150, (1,), generator=g)) for _ in range(world)]
        history.append((full, lens, max(steps)))
        fd, ld = full.to(dev), lens.to(dev)
        outs = []
        for r, x in enumerate(xs):
            lo, hi = shard_bounds(n_total, world, r)
            # the inputs are temporaries of the compute stream, dropped right after the call (record_stream keeps them)
            outs.append(x.step(fd[lo:hi].clone(), ld[lo:hi].clone(), torch.tensor(steps[r], dtype=torch.int32, device=dev)))
        for x, res in zip(xs, outs):
            if it == 0:
                assert res is None
                continue
            x.wait_result()
            tok, ln, st = res
            ef, el, es = history[it - 1]
            assert torch.equal(tok.cpu(), ef) and torch.equal(ln.cpu(), el) and int(st) == es
    for x in xs:
        (tok, ln, st), = x.flush()
        ef, el, es = history[-1]
        assert torch.equal(tok.cpu(), ef) and torch.equal(ln.cpu(), el) and int(st) == es
        assert x.flush() == []
        x.check()


def test_exchange_stage_kick_flush_order(pkg):
    """The serving-loop protocol: stage(i) after decode i, kick() before decode i + 1 hands out result i - 1, flush()
    drains what is left, oldest first."""
    from hmer_img2latex_b200.dist import TokenExchange, shard_bounds
    N = pkg._native
    dev = torch.device("cuda", 0)
    world, n_total, T1 = 2, 6, 9
    nbytes = N.lib().i2l_token_exchange_buffer_bytes(world, n_total, T1)
    bufs = [torch.zeros(nbytes, dtype=torch.uint8, device=dev) for _ in range(world)]
    xs = [TokenExchange(n_total, T1, dev, rank=r, world=world, buffers=bufs) for r in range(world)]
    fulls = [torch.full((n_total, T1), 10 + i, dtype=torch.int64) for i in range(4)]
    lens = torch.arange(1, n_total + 1, dtype=torch.int32)
    got = [[] for _ in xs]
    for i, full in enumerate(fulls):
        for r, x in enumerate(xs):
            res = x.kick()                                    # releases step i - 1, hands out step i - 2
            if res is not None:
                x.wait_result()
                got[r].append(int(res[0][0, 0]))
            lo, hi = shard_bounds(n_total, world, r)
            x.stage(full[lo:hi].to(dev), lens[lo:hi].to(dev), torch.tensor(i + 1, dtype=torch.int32, device=dev))
    for r, x in enumerate(xs):
        # rank 0 drains first: its last read needs rank 1's last write, which rank 1 only releases in ITS flush -- the
        # kick of a flush must therefore not wait for its own read; checked by draining in two passes
        assert x._staged is not None
        res = x.kick()
        x.wait_result()
        got[r].append(int(res[0][0, 0]))
    for r, x in enumerate(xs):
        rest = x.flush()
        got[r] += [int(t[0][0, 0]) for t in rest]
        assert got[r] == [10, 11, 12, 13], got[r]
        x.check()
