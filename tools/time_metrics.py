"""GPU box: time i2l_sequence_metrics on 1024 (prediction, target) pairs of up to 150 tokens, beside the reference's
pure-Python loops restated in oracle/metrics.py on one host core (bounded sample)."""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
import i2l_import
pkg = i2l_import.load()
M = pkg.metrics
B, T = 1024, 151
g = torch.Generator().manual_seed(0)
t = torch.randint(0, 100, (B, T), generator=g)
p = torch.where(torch.rand(B, T, generator=g) < 0.15, torch.randint(0, 100, (B, T), generator=g), t)
lt = torch.randint(100, T + 1, (B,), generator=g, dtype=torch.int32)
lp = (lt - torch.randint(0, 10, (B,), generator=g, dtype=torch.int32)).clamp(min=1)
pc, tc, lpc, ltc = p.cuda(), t.cuda(), lp.cuda(), lt.cuda()
for _ in range(3):
    out = M.sequence_counts(pc, lpc, tc, ltc)
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(20):
    out = M.sequence_counts(pc, lpc, tc, ltc)
e1.record(); torch.cuda.synchronize()
ms = e0.elapsed_time(e1) / 20
cells = float((lp.double() * lt.double()).sum())
print(f"device metrics B={B}: {ms:.4f} ms = {B / ms * 1e3 / 1e6:.2f} M pairs/s, {cells / ms / 1e6:.1f} G DP cells/s")
t0 = time.perf_counter(); res = M.calculate_metrics([p[i, : lp[i]].tolist() for i in range(B)], [t[i, : lt[i]].tolist() for i in range(B)]); t1 = time.perf_counter()
print(f"calculate_metrics from Python lists (pad + H2D + kernel + D2H + float formulas): {(t1 - t0) * 1e3:.1f} ms -> {res}")
sys.path.insert(0, ROOT)
from oracle import metrics as OM
n = 16
t0 = time.perf_counter()
ref = OM.calculate_metrics([p[i, : lp[i]].tolist() for i in range(n)], [t[i, : lt[i]].tolist() for i in range(n)])
t1 = time.perf_counter()
print(f"pure-Python port on one host core, {n} pairs: {(t1 - t0) * 1e3:.1f} ms = {n / (t1 - t0):.1f} pairs/s")
