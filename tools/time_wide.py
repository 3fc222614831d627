"""GPU box: time the greedy loop of a wide decoder (default: the reference's shipped 512 / 512 / 2) on the persistent
whole-GPU kernel (decode_wide.cu) and on the stream-ordered CUDA-graph loop (decoder.streamed = True)."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
import i2l_import
pkg = i2l_import.load()
N = pkg._native
B = int(sys.argv[1]) if len(sys.argv) > 1 else 1024
E, Hd, L = (int(v) for v in sys.argv[2:5]) if len(sys.argv) > 4 else (512, 512, 2)
T = int(sys.argv[5]) if len(sys.argv) > 5 else 150
torch.manual_seed(0)
dec = pkg.LSTMDecoder(512, E, Hd, T, L, 0.0, True, precision="bf16").cuda().eval()
enc = torch.relu(torch.randn(B, E, device="cuda"))
for streamed in (False, True):
    dec.streamed = streamed
    for _ in range(2):
        out = dec.greedy(enc, 1, 2, T, 1.0, N.STOP_NONE)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    n = 5
    e0.record()
    for _ in range(n):
        out = dec.greedy(enc, 1, 2, T, 1.0, N.STOP_NONE)
    e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / n
    print(f"E={E} H={Hd} L={L} B={B} T={T} {'stream-ordered graph loop' if streamed else 'persistent wide kernel'}: "
          f"{ms:.3f} ms, {ms / T * 1e3:.2f} us/step, steps {int(out[2])}", flush=True)
    if not streamed:
        keep = out[0].clone()
    else:
        print("identical tokens:", bool(torch.equal(keep, out[0])))

# sampling loop (Predictor.predict_batch): temperature 0.8, top-k 50, top-p 0.9, fixed uniforms
if len(sys.argv) <= 6 or sys.argv[6] != "nosample":
    u = torch.rand(T, B, device="cuda")
    for streamed in (False, True):
        dec.streamed = streamed
        for _ in range(2):
            out = dec.sample(enc, 1, 2, T, 0.8, 50, 0.9, uniforms=u)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(3):
            out = dec.sample(enc, 1, 2, T, 0.8, 50, 0.9, uniforms=u)
        e1.record(); torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / 3
        n = max(int(out[2]), 1)
        print(f"sampling {'stream-ordered graph loop' if streamed else 'persistent wide kernel'}: {ms:.3f} ms, "
              f"{ms / n * 1e3:.2f} us/step over {n} steps", flush=True)
