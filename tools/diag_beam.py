"""Diagnostic (GPU box): step-0 candidate log-probs of the bf16 persistent beam vs the fp32 oracle."""
import sys, os
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "tests"))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
import torch
import helpers as H
from helpers import oracle
import i2l_import
pkg = i2l_import.load()
sys.path.insert(0, "tests")
from test_gpu_beam_bf16 import run_beam_with_dump

cfg = H.HEADLINE
p = oracle.make_params(cfg, 2, sharp=True)
p["decoder.output_layer.bias"][H.END] += 1.0
m16 = H.build_model(pkg, cfg, p, precision="bf16")
m32 = H.build_model(pkg, cfg, p, precision="fp32")
B, K, T = 6, 5, 8
x = H.make_images(cfg, B)
enc_ref = oracle.encoder(p, x, cfg)
enc = m32.encoder(x.cuda())
out, olen, score, trp, trt, trs, ctok, clogp = run_beam_with_dump(pkg, m16, enc, T, K)
for b in range(3):
    seq, sc, trace, cands = oracle.beam_search(p, enc_ref[b:b + 1], H.START, H.END, T, K, cfg, return_cands=True)
    for t in range(3):
        print("img", b, "step", t)
        print("  oracle kept:", [(pb, tk, round(s, 4)) for pb, tk, s in trace[t]])
        print("  device kept:", list(zip(trp[t, b].tolist(), trt[t, b].tolist(), [round(v, 4) for v in trs[t, b].tolist()])))
        print("  device cand slot0:", list(zip(ctok[t, b, 0].tolist(), [round(v, 4) for v in clogp[t, b, 0].tolist()])))
