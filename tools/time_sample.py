"""GPU box: time the sampling loop (general path) for the headline decoder, fp32 vs bf16 GEMMs."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
import i2l_import
pkg = i2l_import.load()
B, T = 1024, 150
for prec in ("bf16", "fp32"):
    torch.manual_seed(0)
    m = pkg.Seq2SeqModel("cnn_lstm", 512, dict(img_height=64, img_width=320, channels=3, embedding_dim=256),
                         dict(hidden_dim=256, lstm_layers=1, attention=True), precision=prec).cuda().eval()
    enc = torch.relu(torch.randn(B, 256)).cuda()
    for _ in range(2):
        m.decoder.sample(enc, 1, 2, T, temperature=0.8, top_k=50, top_p=0.9, seed=1)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(3):
        m.decoder.sample(enc, 1, 2, T, temperature=0.8, top_k=50, top_p=0.9, seed=1)
    e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / 3
    print(f"sample loop {prec}: {ms:.3f} ms  {ms / T * 1e3:.1f} us/step")
