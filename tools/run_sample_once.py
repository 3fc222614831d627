"""GPU box helper for ncu: one general-path sampling decode (B=1024, few steps)."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
import i2l_import
pkg = i2l_import.load()
T = int(sys.argv[1]) if len(sys.argv) > 1 else 3
torch.manual_seed(0)
m = pkg.Seq2SeqModel("cnn_lstm", 512, dict(img_height=64, img_width=320, channels=3, embedding_dim=256),
                     dict(hidden_dim=256, lstm_layers=1, attention=True), precision="bf16").cuda().eval()
enc = torch.relu(torch.randn(1024, 256)).cuda()
m.decoder.sample(enc, 1, 2, T, temperature=0.8, top_k=50, top_p=0.9, seed=1)
torch.cuda.synchronize()
