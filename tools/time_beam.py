"""Diagnostic (GPU box): per-phase clock64 stamps of one step of the persistent beam kernel.
Usage: python tools/time_beam.py [step] [B] [K]"""
import ctypes as C, os, sys
# needs the diagnostics library: make -C hmer-img2latex_b200/csrc diag
os.environ.setdefault("I2L_LIB", os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "hmer-img2latex_b200", "csrc", "libi2l_b200_diag.so"))
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
import i2l_import
pkg = i2l_import.load()
step = int(sys.argv[1]) if len(sys.argv) > 1 else 20
B = int(sys.argv[2]) if len(sys.argv) > 2 else 512
K = int(sys.argv[3]) if len(sys.argv) > 3 else 5
torch.manual_seed(0)
m = pkg.Seq2SeqModel("cnn_lstm", 512, dict(img_height=64, img_width=320, channels=3, embedding_dim=256),
                     dict(hidden_dim=256, lstm_layers=1, attention=True), precision="bf16").cuda().eval()
enc = torch.relu(torch.randn(B, 256)).cuda()
lib = pkg._native.lib()
lib.i2l_debug_set_beam_ts.argtypes = [C.c_void_p, C.c_int32]; lib.i2l_debug_set_beam_ts.restype = C.c_int
ts = torch.zeros(64, dtype=torch.int64, device="cuda")
NAMES = {0: "step start", 1: "gtok loads issued", 2: "GDONE waited", 3: "gate tmem ld done", 4: "cell update + h stored",
         5: "bar + bulk issued", 6: "LDONE waited", 7: "logits -> LT written", 8: "bar", 9: "insertion done",
         10: "8-lane merges done", 11: "sumexp + record stored", 12: "bar + bulk sent + retire", 18: "TOK waited",
         13: "4-way merge + lse done", 14: "cand keys written", 19: "bar", 20: "rank counting done", 15: "bar",
         16: "bookkeeping + pub written", 17: "bar (end of step)", 21: "exp loop done", 22: "record stored", 23: "bar", 24: "bulk issued",
         25: "lse + 4-way ranks done"}
m.decoder.beam(enc, 1, 2, 150, K)
lib.i2l_debug_set_beam_ts(ts.data_ptr(), step)
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record(); m.decoder.beam(enc, 1, 2, 150, K); e1.record()
torch.cuda.synchronize()
lib.i2l_debug_set_beam_ts(None, 0)
t = ts.cpu().tolist()
print(f"beam call B={B} K={K}: {e0.elapsed_time(e1):.3f} ms")
order = [0, 1, 2, 3, 4, 5, 6, 7, 8, 9, 10, 11, 23, 24, 12, 18, 25, 13, 14, 19, 20, 15, 16, 17]
prev = t[0]
for i in order:
    print(f"  ts[{i:2d}] {NAMES[i]:28s} {t[i] - t[0]:7d} cyc  (+{t[i] - prev})")
    prev = t[i]
