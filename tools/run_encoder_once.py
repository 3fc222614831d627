import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import torch, i2l_import, helpers as H
from helpers import oracle
pkg = i2l_import.load()
B = int(sys.argv[1]) if len(sys.argv) > 1 else 2
p = oracle.make_params(H.HEADLINE, 1)
m = H.build_model(pkg, H.HEADLINE, p, "bf16")
x = H.make_images(H.HEADLINE, B)
out = m.encoder(x.cuda()); torch.cuda.synchronize()
print("rel err", H.rel_err(out, oracle.cnn_encoder(p, x)))
