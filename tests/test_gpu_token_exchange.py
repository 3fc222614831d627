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
