"""GPU box: time the decode loops of an arbitrary decoder shape (greedy + sampling), us per step.
Usage: python tools/time_decoder.py E H L V [B] [T]     e.g. 512 512 2 512  (configs/config.yaml:45-50 of the reference)"""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
import i2l_import
pkg = i2l_import.load()
E, H, L, V = (int(a) for a in sys.argv[1:5])
B = int(sys.argv[5]) if len(sys.argv) > 5 else 1024
T = int(sys.argv[6]) if len(sys.argv) > 6 else 150
torch.manual_seed(0)
for prec in ("bf16", "fp32"):
    dec = pkg.LSTMDecoder(V, E, H, T, L, 0.0, True, precision=prec).cuda().eval()
    enc = torch.relu(torch.randn(B, E)).cuda()
    for name, fn in (("greedy", lambda: dec.greedy(enc, 1, 2, T, 1.0, pkg._native.STOP_NONE)),
                     ("sample", lambda: dec.sample(enc, 1, 2, T, 0.8, 50, 0.9, seed=1))):
        for _ in range(3):
            fn()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        reps = 5
        e0.record()
        for _ in range(reps):
            fn()
        e1.record(); torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / reps
        flops = B * (2 * (2 * E + H) * 4 * H + (L - 1) * 2 * 2 * H * 4 * H + 2 * H * V)
        print(f"E={E} H={H} L={L} V={V} B={B} {prec} {name}: {ms:.3f} ms, {ms / T * 1e3:.2f} us/step, "
              f"{flops / (ms / T * 1e-3) / 1e12:.1f} TFLOP/s (SURVEY 8a4 live FLOPs)")
