"""GPU box: time the ResNet encoders (bf16 tcgen05 path) with the per-kernel CUDA-event breakdown.
Usage: python tools/bench_resnet.py [resnet18|resnet50] [B] [W] [precision]"""
import os, sys, json
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
import i2l_import
pkg = i2l_import.load()
N = pkg._native
name = sys.argv[1] if len(sys.argv) > 1 else "resnet18"
B = int(sys.argv[2]) if len(sys.argv) > 2 else 1024
W = int(sys.argv[3]) if len(sys.argv) > 3 else 320
prec = sys.argv[4] if len(sys.argv) > 4 else "bf16"
MFLOP_PER_COL = {"resnet18": 4.627, "resnet50": 10.43}[name]      # SURVEY 8d, per image per pixel column
torch.manual_seed(0)
enc = pkg.ResNetEncoder(64, W, 3, name, 256, precision=prec).cuda().eval()
x = [torch.randn(B, 3, 64, W, device="cuda") for _ in range(2)]
lib = N.lib()
with torch.no_grad():
    for i in range(3):
        enc(x[i & 1])
    torch.cuda.synchronize()
    lib.i2l_prof_reset(); lib.i2l_prof_enable(1)
    reps = 5
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for i in range(reps):
        enc(x[i & 1])
    e1.record(); torch.cuda.synchronize()
    lib.i2l_prof_enable(0)
ms = e0.elapsed_time(e1) / reps
fl = MFLOP_PER_COL * 1e6 * W * B
prof = N.prof_results()
print(json.dumps({"model": name, "B": B, "W": W, "precision": prec, "ms": round(ms, 3), "images_per_s": round(B / ms * 1e3, 1),
                  "tflops": round(fl / ms / 1e9, 1),
                  "kernels_ms": {k: [v[0] // reps, round(v[1] / reps, 4)] for k, v in sorted(prof.items())}}))
