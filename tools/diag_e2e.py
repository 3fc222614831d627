"""Diagnostic (GPU box): where does the host-buffer (e2e) path spend its time per batch?"""
import sys, os, time
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
import torch
import i2l_import
pkg = i2l_import.load()
N = pkg._native
dev = torch.device("cuda", 0)
torch.manual_seed(0)
model = pkg.Seq2SeqModel("cnn_lstm", 512, dict(img_height=64, img_width=320, channels=3, embedding_dim=256),
                         dict(hidden_dim=256, lstm_layers=1, attention=True), precision="bf16").to(dev).eval()
B = 1024
xm = torch.randn(B, 3, 64, 320)
hosts = {"fp32": xm.pin_memory(), "bf16": xm.bfloat16().pin_memory(),
         "uint8": ((xm.clamp(-1, 1) + 1) * 127.5).round().to(torch.uint8).pin_memory()}

def ev():
    return torch.cuda.Event(enable_timing=True)

with torch.no_grad():
    for nm, h in hosts.items():
        d = torch.empty_like(h, device=dev)
        for _ in range(3):
            d.copy_(h, non_blocking=True)
        torch.cuda.synchronize()
        a, b = ev(), ev()
        a.record()
        for _ in range(10):
            d.copy_(h, non_blocking=True)
        b.record(); torch.cuda.synchronize()
        ms = a.elapsed_time(b) / 10
        print(f"H2D {nm}: {ms:.3f} ms  ({h.numel() * h.element_size() / ms / 1e6:.1f} GB/s)")
    d8 = hosts["uint8"].to(dev)
    for _ in range(3):
        y = pkg.normalize_u8(d8, "pm1", out_dtype=torch.bfloat16)
    a, b = ev(), ev(); a.record()
    for _ in range(10):
        y = pkg.normalize_u8(d8, "pm1", out_dtype=torch.bfloat16)
    b.record(); torch.cuda.synchronize()
    print(f"normalize_u8 -> bf16: {a.elapsed_time(b) / 10:.3f} ms")
    for nm, h in hosts.items():
        for steps in (10, 40):
            def gen(k):
                for _ in range(k):
                    yield h
            for _ in model.greedy_stream(gen(3), 1, 2, 150):
                pass
            torch.cuda.synchronize()
            t0 = time.perf_counter()
            for _ in model.greedy_stream(gen(steps), 1, 2, 150):
                pass
            torch.cuda.synchronize()
            dt = time.perf_counter() - t0
            print(f"e2e {nm} steps={steps}: {dt / steps * 1e3:.3f} ms/batch  {B * steps / dt:.0f} img/s")
