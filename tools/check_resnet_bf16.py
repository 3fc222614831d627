"""Diagnostic (GPU box): bf16 tcgen05 ResNet encoder vs. the fp32 CPU oracle and the fp32 CUDA path.
Usage: python tools/check_resnet_bf16.py [resnet18|resnet50] [B] [W]"""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import torch
import i2l_import
import helpers as H
from helpers import oracle
pkg = i2l_import.load()
name = sys.argv[1] if len(sys.argv) > 1 else "resnet18"
B = int(sys.argv[2]) if len(sys.argv) > 2 else 3
W = int(sys.argv[3]) if len(sys.argv) > 3 else 128
cfg = dict(H.R18 if name == "resnet18" else H.R50); cfg["model_name"] = name
p = oracle.make_params(cfg, 0)
x = H.make_images(cfg, B, width=W)
ref = oracle.resnet_encoder(p, x, name)
m16 = H.build_model(pkg, cfg, p, "bf16")
out = m16.encoder(x.cuda()); torch.cuda.synchronize()
m32 = H.build_model(pkg, cfg, p, "fp32")
o32 = m32.encoder(x.cuda()); torch.cuda.synchronize()
print(f"{name} B={B} W={W}: bf16 rel err vs oracle {H.rel_err(out, ref):.3e}   fp32 rel err {H.rel_err(o32, ref):.3e}   max|ref| {float(ref.abs().max()):.3f}")
print("ref ", ref[0, :6].tolist()); print("bf16", out[0, :6].cpu().tolist())
