"""CPU suite: the N>1 host logic (batch sharding + token all-gather + stop-step reduction)
on a world_size-2 gloo group."""
import os
import sys

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _worker(rank, world, port, n_total, T1, q):
    sys.path.insert(0, ROOT)
    import i2l_import
    i2l_import.load()
    from hmer_img2latex_b200.dist import gather_tokens, shard_bounds
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    lo, hi = shard_bounds(n_total, world, rank)
    full = torch.arange(n_total * T1, dtype=torch.int64).reshape(n_total, T1)
    lens = torch.arange(n_total, dtype=torch.int32) % T1
    steps = torch.tensor(3 + 5 * rank, dtype=torch.int32)
    tok, ln, st = gather_tokens(full[lo:hi].clone(), lens[lo:hi].clone(), steps, n_total)
    ok = torch.equal(tok, full) and torch.equal(ln, lens) and int(st) == 3 + 5 * (world - 1)
    q.put((rank, bool(ok)))
    dist.destroy_process_group()


@pytest.mark.parametrize("n_total", [7, 8, 1])
def test_gather_tokens_world2(n_total):
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29500 + (os.getpid() + n_total) % 500
    procs = [ctx.Process(target=_worker, args=(r, 2, port, n_total, 6, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = [q.get(timeout=120) for _ in procs]
    for p in procs:
        p.join(timeout=60)
    assert sorted(res) == [(0, True), (1, True)]


def _eval_worker(rank, world, port, q):
    """Sharded evaluate / validation on gloo: per-shard counts (here from the CPU oracle -- the device kernels produce
    the same integers, tests/test_gpu_metrics.py) gathered in image order give the single-process means bit for bit;
    shard-mean losses combine to the global token mean."""
    sys.path.insert(0, ROOT)
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    import i2l_import
    pkg = i2l_import.load()
    from hmer_img2latex_b200.dist import gather_counts, reduce_validation, shard_bounds
    from oracle import metrics as OM
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    g = torch.Generator().manual_seed(5)
    n = 11
    preds = [torch.randint(0, 6, (int(torch.randint(1, 30, (1,), generator=g)),), generator=g).tolist() for _ in range(n)]
    tgts = [torch.randint(0, 6, (int(torch.randint(1, 30, (1,), generator=g)),), generator=g).tolist() for _ in range(n)]
    lo, hi = shard_bounds(n, world, rank)
    local = torch.tensor([[OM.edit_distance(p, t)] + [OM.clipped_matches(p, t, k) for k in range(1, 5)] + [len(p), len(t), 0]
                          for p, t in zip(preds[lo:hi], tgts[lo:hi])], dtype=torch.int32).reshape(hi - lo, 8)
    full = gather_counts(local, n)
    both = [pkg.metrics.scores_from_counts(r, 4) for r in full.tolist()]
    got = {"bleu": sum(b for _, b in both) / n, "levenshtein": sum(l for l, _ in both) / n, "batch_size": n}
    ok = got == OM.calculate_metrics(preds, tgts)
    # validation: shard-mean losses -> global token mean
    logits = torch.randn(n, 9, 13, generator=g) * 3
    tg = torch.randint(0, 13, (n, 9), generator=g)
    ref_loss, rc, rt = OM.validation_loss_accuracy(logits, tg, 0, 0.1)
    l, c, t = OM.validation_loss_accuracy(logits[lo:hi], tg[lo:hi], 0, 0.1)
    gl, gc, gt = reduce_validation(l, c, t)
    ok = ok and (gc, gt) == (rc, rt) and abs(gl - float(ref_loss)) < 1e-6 * abs(float(ref_loss))
    q.put((rank, bool(ok)))
    dist.destroy_process_group()


def test_sharded_evaluate_world2():
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29500 + (os.getpid() + 77) % 500
    procs = [ctx.Process(target=_eval_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = [q.get(timeout=180) for _ in procs]
    for p in procs:
        p.join(timeout=60)
    assert sorted(res) == [(0, True), (1, True)]


def test_shard_bounds_cover(pkg):
    from hmer_img2latex_b200.dist import shard_bounds
    for n in (0, 1, 5, 8, 1024, 1025):
        for ws in (1, 2, 3, 8):
            cuts = [shard_bounds(n, ws, r) for r in range(ws)]
            assert cuts[0][0] == 0 and cuts[-1][1] == n
            assert all(a[1] == b[0] for a, b in zip(cuts, cuts[1:]))
            sizes = [b - a for a, b in cuts]
            assert max(sizes) - min(sizes) <= 1
