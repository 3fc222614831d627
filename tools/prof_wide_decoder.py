"""GPU box, diagnostics library: per-kernel device time of the graph-replayed general decode loop (event nodes inside
the captured graph).  Usage: python tools/prof_wide_decoder.py E H L V [B] [T]"""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
os.environ.setdefault("I2L_LIB", os.path.join(ROOT, "hmer-img2latex_b200", "csrc", "libi2l_b200_diag.so"))
sys.path.insert(0, ROOT)
import torch
import i2l_import
pkg = i2l_import.load()
N = pkg._native
E, H, L, V = (int(a) for a in sys.argv[1:5])
B = int(sys.argv[5]) if len(sys.argv) > 5 else 1024
T = int(sys.argv[6]) if len(sys.argv) > 6 else 150
dec = pkg.LSTMDecoder(V, E, H, T, L, 0.0, True, precision="bf16").cuda().eval()
enc = torch.relu(torch.randn(B, E)).cuda()
lib = N.lib()
lib.i2l_prof_reset(); lib.i2l_prof_enable(1)
for _ in range(3):
    dec.greedy(enc, 1, 2, T, 1.0, N.STOP_NONE)
torch.cuda.synchronize()
lib.i2l_prof_enable(0)
for k, (cnt, ms) in sorted(N.prof_results().items()):
    print(f"{k:28s} launches {cnt:5d}  total {ms:8.3f} ms  -> {ms / cnt * 1e3:7.2f} us each")
