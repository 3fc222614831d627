"""GPU box: time the sampling decode loop (headline decoder, B=1024, 150 steps): persistent kernel vs the
stream-ordered general path (decoder.streamed = True, I2L_BF16_STREAMED)."""
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
for label, env in (("persistent", None), ("general", "1")):
    m.decoder.streamed = bool(env)
    for args in ((0.8, 50, 0.9), (1.0, 0, 0.9), (1.0, 200, 0.0), (0.7, 0, 0.0)):
        for _ in range(2):
            m.decoder.sample(enc, 1, 2, T, *args, seed=1)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(5):
            tok, ln, st = m.decoder.sample(enc, 1, 2, T, *args, seed=1)
        e1.record(); torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / 5
        n = int(st)
        print(f"{label:10s} T={args[0]} k={args[1]} p={args[2]}: {ms:.3f} ms, {n} steps, {ms / max(n, 1) * 1e3:.2f} us/step")
