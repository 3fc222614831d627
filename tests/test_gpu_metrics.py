"""Evaluation metrics on the device through the C-ABI (`i2l_sequence_metrics`) against the
live-reference golden vectors (training/metrics.py) and the CPU oracle.  Integer counts are bit-exact and the
host applies the reference's float formulas with the same libm calls, so the scores are compared with ==."""
import os

import numpy as np
import pytest
import torch

import helpers as H
from helpers import oracle

pytestmark = pytest.mark.gpu

G = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
OM = oracle.metrics


def unpad(rows):
    return [[int(t) for t in r if t >= 0] for r in rows]


def test_metrics_vs_reference_golden(pkg):
    d = np.load(os.path.join(G, "metrics.npz"))
    preds, tgts = unpad(d["pred"]), unpad(d["tgt"])
    M = pkg.metrics
    assert [M.levenshtein_distance(p, t) for p, t in zip(preds[:12], tgts[:12])] == d["lev"][:12].tolist()
    assert [M.bleu_n_score(p, t, 4) for p, t in zip(preds[:12], tgts[:12])] == d["bleu4"][:12].tolist()
    assert [M.bleu_n_score(p, t, 2) for p, t in zip(preds, tgts)] == d["bleu2"].tolist()
    res = M.calculate_metrics(preds[2:], tgts[2:])
    assert [res["bleu"], res["levenshtein"], float(res["batch_size"])] == d["mean"].tolist()
    # whole set in one launch, all scores
    p, pl = M._pad(preds); t, tl = M._pad(tgts)
    rows = M.sequence_counts(p.cuda(), pl.cuda(), t.cuda(), tl.cuda()).tolist()
    assert [M.scores_from_counts(r)[0] for r in rows] == d["lev"].tolist()
    assert [M.scores_from_counts(r)[1] for r in rows] == d["bleu4"].tolist()


@pytest.mark.parametrize("V,maxlen,B", [(3, 40, 64), (50, 150, 96), (512, 151, 64), (4, 700, 12), (2, 9, 200)])
def test_counts_vs_oracle(pkg, V, maxlen, B):
    rng = np.random.default_rng(V * 1000 + maxlen)
    preds, tgts = [], []
    for i in range(B):
        t = rng.integers(0, V, size=int(rng.integers(0, maxlen + 1))).tolist()
        if i % 3 == 0:
            p = [x for x in t if rng.random() > 0.1]
        elif i % 3 == 1:
            p = rng.integers(0, V, size=int(rng.integers(0, maxlen + 1))).tolist()
        else:
            p = list(t)
        preds.append(p); tgts.append(t)
    M = pkg.metrics
    p, pl = M._pad(preds); t, tl = M._pad(tgts)
    rows = M.sequence_counts(p.cuda(), pl.cuda(), t.cuda(), tl.cuda()).tolist()
    for r, a, b in zip(rows, preds, tgts):
        assert r[0] == OM.edit_distance(a, b), (len(a), len(b))
        assert r[1:5] == [OM.clipped_matches(a, b, g) for g in range(1, 5)]
        assert r[5:7] == [len(a), len(b)]
        assert M.scores_from_counts(r) == (OM.levenshtein_distance(a, b), OM.bleu_n_score(a, b, 4))


def test_metrics_properties_full_size(pkg):
    """BASELINE-size batch (1024 decoded sequences, T = 150): distance properties that need no oracle --
    d(a,a) = 0 with all n-grams matched, symmetry, |len(a)-len(b)| <= d <= max(len), d(a, a+suffix) = len(suffix)."""
    g = torch.Generator().manual_seed(0)
    B, T = 1024, 151
    a = torch.randint(0, 6, (B, T), generator=g)
    b = torch.randint(0, 6, (B, T), generator=g)
    la = torch.randint(0, T + 1, (B,), generator=g, dtype=torch.int32)
    lb = torch.randint(0, T + 1, (B,), generator=g, dtype=torch.int32)
    M = pkg.metrics
    ac, bc, lac, lbc = a.cuda(), b.cuda(), la.cuda(), lb.cuda()
    same = M.sequence_counts(ac, lac, ac, lac).cpu()
    assert (same[:, 0] == 0).all()
    for n in range(1, 5):
        assert torch.equal(same[:, n], (la - n + 1).clamp(min=0))
    ab, ba = M.sequence_counts(ac, lac, bc, lbc).cpu(), M.sequence_counts(bc, lbc, ac, lac).cpu()
    assert torch.equal(ab[:, 0], ba[:, 0])
    assert (ab[:, 0] >= (la - lb).abs()).all() and (ab[:, 0] <= torch.maximum(la, lb)).all()
    short = (la // 2).to(torch.int32)
    pre = M.sequence_counts(ac, short.cuda(), ac, lac).cpu()                 # a[:k] vs a: k insertions
    assert torch.equal(pre[:, 0], la - short)
    for i in (0, 17, 1023):
        x, y = a[i, : la[i]].tolist(), b[i, : lb[i]].tolist()
        assert ab[i, 0] == OM.edit_distance(x, y) and ab[i, 1:5].tolist() == [OM.clipped_matches(x, y, n) for n in range(1, 5)]


def test_metrics_edge_cases(pkg):
    M = pkg.metrics
    assert M.levenshtein_distance([], []) == 1.0 and M.bleu_n_score([], [1]) == 0.0
    assert M.levenshtein_distance([1, 2], []) == 0.0
    with pytest.raises(ZeroDivisionError):
        M.calculate_metrics([], [])
    with pytest.raises(AssertionError):
        M.calculate_metrics([[1]], [[1], [2]])
    big = 2 ** 40
    assert M.levenshtein_distance([big, big + 1], [big, big + 2]) == 0.5      # ids compare as int64
    with pytest.raises(RuntimeError, match="too long"):
        M.sequence_counts(torch.zeros(1, 20000, dtype=torch.long).cuda(), torch.tensor([5]).cuda(),
                          torch.zeros(1, 20000, dtype=torch.long).cuda(), torch.tensor([5]).cuda())


def test_filter_ids(pkg):
    M = pkg.metrics
    g = torch.Generator().manual_seed(1)
    ids = torch.randint(0, 9, (37, 70), generator=g)
    ln = torch.randint(0, 71, (37,), generator=g, dtype=torch.int32)
    out, ol = M.filter_ids(ids.cuda(), ln.cuda(), [0, 1, 2, 3])
    for r in range(37):
        ref = [x for x in ids[r, : ln[r]].tolist() if x > 3]
        assert out[r, : ol[r]].tolist() == ref
    out, ol = M.filter_ids(ids.cuda(), None, [5])
    assert out[3, : ol[3]].tolist() == [x for x in ids[3].tolist() if x != 5]


def test_evaluate_batch_equals_string_route(pkg):
    """Predictor.evaluate_batch (device-resident ids) == the reference's evaluate loop body (cli.py:448-495):
    predict_batch -> strings -> tokenizer.encode -> calculate_metrics."""
    cfg = H.SMALL
    p = oracle.make_params(cfg, 1, sharp=True)
    m = H.build_model(pkg, cfg, p)
    tok = pkg.LaTeXTokenizer(); tok.default_init()
    pred = pkg.Predictor(m, tok)
    B, T = 24, 30
    x = H.make_images(cfg, B).abs().clamp(max=1.0)
    g = torch.Generator().manual_seed(4)
    targets = torch.randint(4, 46, (B, T), generator=g)
    targets[:, 0] = tok.start_token_id
    for b in range(B):
        e = int(torch.randint(2, T, (1,), generator=g))
        targets[b, e] = tok.end_token_id
        targets[b, e + 1:] = tok.pad_token_id
    strs = pred.predict_batch(list(x), max_length=25, batch_size=B)
    preds = [tok.encode(s) for s in strs]
    tg = [[i for i in row if i != tok.pad_token_id] for row in targets.tolist()]
    ref = OM.calculate_metrics(preds, tg)
    got = pred.evaluate_batch(list(x), targets.cuda(), max_length=25, batch_size=16)
    assert got == ref
    assert pkg.metrics.calculate_metrics(preds, tg) == ref
