"""Debug aid: dumps gates / h / logits of one step of the persistent decode kernel and compares
them with the CPU oracle.  Usage (GPU box): python tools/debug_persistent.py [step]"""
import ctypes as C
import os
import sys
# needs the diagnostics library: make -C hmer-img2latex_b200/csrc diag
os.environ.setdefault("I2L_LIB", os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "hmer-img2latex_b200", "csrc", "libi2l_b200_diag.so"))

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import i2l_import
import helpers as H
from helpers import oracle

pkg = i2l_import.load()
step = int(sys.argv[1]) if len(sys.argv) > 1 else 0
cfg = H.HEADLINE
p = oracle.make_params(cfg, 1, sharp=True)
B = int(sys.argv[3]) if len(sys.argv) > 3 else 32
T = step + 2
m32 = H.build_model(pkg, cfg, p, "fp32"); m16 = H.build_model(pkg, cfg, p, "bf16")
x = H.make_images(cfg, B)
enc_ref = oracle.encoder(p, x, cfg)
lib = pkg._native.lib()
dump = int(sys.argv[2]) if len(sys.argv) > 2 else 1
dbg = torch.zeros(16 + 400000 + 200000, device="cuda"); dbg[0] = step; dbg[1] = dump
lib.i2l_debug_set_buffer.argtypes = [C.c_void_p]; lib.i2l_debug_set_buffer.restype = C.c_int
assert lib.i2l_debug_set_buffer(C.c_void_p(dbg.data_ptr())) == 0
enc = m32.encoder(x.cuda())
tokens, lengths, steps = m16.decoder.greedy(enc, H.START, H.END, T)
torch.cuda.synchronize()
d = dbg.cpu()[16:]
NAMES = {0: 'epiG start', 1: 'gtok loads issued', 2: 'GDONE waited', 3: 'tmem ld done', 4: 'pointwise+h st done', 5: 'bar+bulk issued', 6: 'LDONE waited', 7: 'argmax partial', 8: 'bar', 9: 'xchg sent', 10: 'TOK waited', 11: 'tok final', 12: 'bar (end step)', 13: 'xchg: combine done', 14: 'xchg: expect_tx done', 15: 'xchg: first st.async', 16: 'mma: before HFULL wait', 17: 'mma: HFULL ok', 18: 'mma: L issued', 19: 'mma: G issued'}
if not dump:
    ts = dbg.cpu()[16 + 300000: 16 + 300000 + 64].view(torch.int64)
    t0 = int(ts[0])
    for k in sorted(NAMES, key=lambda k: int(ts[k])):
        print(f'  ts[{k:2d}] {NAMES[k]:28s} {int(ts[k]) - t0:8d} cyc')
    sys.exit(0)
# oracle up to `step`
tok = torch.full((B, 1), H.START, dtype=torch.long); hid = None
for s in range(step):
    out, hid = oracle.decode_step(p, enc_ref, tok, hid, cfg); tok = out.squeeze(1).argmax(-1, keepdim=True)
Hd = 256
h0 = hid[0][0] if hid else torch.zeros(B, Hd); c0 = hid[1][0] if hid else torch.zeros(B, Hd)
emb = p["decoder.embedding.weight"][tok.squeeze(1)]
xin = torch.cat([emb, enc_ref], 1)
gates = xin @ p["decoder.lstm.weight_ih_l0"].T + p["decoder.lstm.bias_ih_l0"] + h0 @ p["decoder.lstm.weight_hh_l0"].T + p["decoder.lstm.bias_hh_l0"]
mm = h0 @ p["decoder.lstm.weight_hh_l0"].T
out, (h1, c1) = oracle.decode_step(p, enc_ref, tok, hid, cfg)
logits = out.squeeze(1)
# kernel dumps
ga = d[:32768].reshape(4, 2, 128, 32); gm = d[200000:200000 + 32768].reshape(4, 2, 128, 32)
ref_g = torch.zeros(4, 2, 128, 32); ref_m = torch.zeros(4, 2, 128, 32)
for r in range(4):
    for t in range(2):
        for pp in range(128):
            q, l = pp // 32, pp % 32
            unit = 64 * r + 16 * q + (l % 16); gate = 2 * t + (1 if l >= 16 else 0)
            ref_g[r, t, pp] = gates[:, gate * 256 + unit]; ref_m[r, t, pp] = mm[:, gate * 256 + unit]
print("gates   max|diff|", float((ga - ref_g).abs().max()), "ref max", float(ref_g.abs().max()))
print("mma-G   max|diff|", float((gm - ref_m).abs().max()), "ref max", float(ref_m.abs().max()))
hk = d[65536:65536 + 8192].reshape(32, 256)
print("h       max|diff|", float((hk - h1[0]).abs().max()), "ref max", float(h1.abs().max()))
lk = d[131072:131072 + 16384].reshape(4, 128, 32)
ref_l = logits.T.reshape(4, 128, 32)
print("logits  max|diff|", float((lk - ref_l).abs().max()), "ref max", float(ref_l.abs().max()))
print("kernel argmax", lk.reshape(512, 32).argmax(0)[:8].tolist(), "oracle", logits.argmax(1)[:8].tolist(), "tokens", tokens[:8, step + 1].tolist())
for r in range(4):
    print(" rank", r, "logit diff", float((lk[r] - ref_l[r]).abs().max()), "gate diff", float((ga[r] - ref_g[r]).abs().max()))
