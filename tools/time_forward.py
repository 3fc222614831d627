"""GPU box: time the teacher-forced decoder pass (i2l_decoder_forward), headline decoder, B x T tokens."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
import i2l_import
pkg = i2l_import.load()
B = int(sys.argv[1]) if len(sys.argv) > 1 else 1024
T = int(sys.argv[2]) if len(sys.argv) > 2 else 150
torch.manual_seed(0)
for prec in ("bf16", "fp32"):
    m = pkg.Seq2SeqModel("cnn_lstm", 512, dict(img_height=64, img_width=320, channels=3, embedding_dim=256),
                         dict(hidden_dim=256, lstm_layers=1, attention=True), precision=prec).cuda().eval()
    enc = torch.relu(torch.randn(B, 256)).cuda()
    tgt = torch.randint(0, 512, (B, T)).cuda()
    for _ in range(2):
        m.decoder(enc, tgt)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    n = 5
    e0.record()
    for _ in range(n):
        m.decoder(enc, tgt)
    e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / n
    print(f"teacher-forced forward {prec} B={B} T={T}: {ms:.3f} ms  {ms / T * 1e3:.2f} us/step  {B * T / ms * 1e3 / 1e6:.1f} M tokens/s")
