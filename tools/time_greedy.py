"""GPU box: time the persistent greedy kernel alone (B=1024, 150 steps).  Honors I2L_LIB (ablation builds)."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
import i2l_import
pkg = i2l_import.load()
B = int(sys.argv[1]) if len(sys.argv) > 1 else 1024
T = int(sys.argv[2]) if len(sys.argv) > 2 else 150
torch.manual_seed(0)
m = pkg.Seq2SeqModel("cnn_lstm", 512, dict(img_height=64, img_width=320, channels=3, embedding_dim=256),
                     dict(hidden_dim=256, lstm_layers=1, attention=True), precision="bf16").cuda().eval()
enc = torch.relu(torch.randn(B, 256)).cuda()
lib = pkg._native.lib()
for _ in range(3):
    m.decoder.greedy(enc, 1, 2, T, 1.0, 0)
torch.cuda.synchronize()
lib.i2l_prof_reset(); lib.i2l_prof_enable(1)
for _ in range(10):
    m.decoder.greedy(enc, 1, 2, T, 1.0, 0)
torch.cuda.synchronize()
lib.i2l_prof_enable(0)
pr = pkg._native.prof_results()
c, ms = pr["dec.greedy_persistent"]
print(f"{os.environ.get('I2L_LIB', 'default'):40s} greedy B={B} T={T}: {ms / c:.4f} ms  {ms / c / T * 1e3:.3f} us/step")
