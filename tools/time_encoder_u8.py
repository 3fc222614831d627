"""GPU box: time CNNEncoder.forward_u8 per kernel (i2l_prof_*), headline 3x64x320 (or C H W from argv), B images.
I2L_LIB=.../libi2l_b200_diag.so + I2L_CONV1_NG / I2L_CONV1_U8_LEGACY select the A-B variants of conv1."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
import i2l_import
pkg = i2l_import.load()
N = pkg._native
B = int(sys.argv[1]) if len(sys.argv) > 1 else 1024
C_, H_, W_ = (int(v) for v in sys.argv[2:5]) if len(sys.argv) > 4 else (3, 64, 320)
torch.manual_seed(0)
m = pkg.Seq2SeqModel("cnn_lstm", 512, dict(img_height=H_, img_width=W_, channels=C_, embedding_dim=256),
                     dict(hidden_dim=256, lstm_layers=1, attention=True), precision="bf16").cuda().eval()
xs = [torch.randint(0, 256, (B, C_, H_, W_), dtype=torch.uint8, device="cuda") for _ in range(3)]
lib = N.lib()
with torch.no_grad():
    for i in range(5):
        m.encoder.forward_u8(xs[i % 3])
    torch.cuda.synchronize()
    lib.i2l_prof_reset(); lib.i2l_prof_enable(1)
    n = 20
    for i in range(n):
        m.encoder.forward_u8(xs[i % 3])
    torch.cuda.synchronize()
    lib.i2l_prof_enable(0)
print(os.environ.get("I2L_CONV1_NG"), os.environ.get("I2L_CONV1_U8_LEGACY"),
      {k: round(v[1] / v[0], 4) for k, v in sorted(N.prof_results().items())})
