"""GPU box: time the device-side image preparation (i2l_resize_pad_u8) on a ragged batch of formula-sized
grey images -> 64 x W, beside Pillow on the host cores (same images, ResizeWithAspectRatio arithmetic)."""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np
import torch
import i2l_import
pkg = i2l_import.load()
P = pkg.preprocess
B = int(sys.argv[1]) if len(sys.argv) > 1 else 1024
TW = int(sys.argv[2]) if len(sys.argv) > 2 else 800
rng = np.random.default_rng(0)
imgs = []
for i in range(B):
    h = int(rng.integers(32, 128)); w = int(h * rng.uniform(2.0, 14.0))
    imgs.append(rng.integers(0, 256, size=(h, w), dtype=np.uint8))
src_bytes = sum(a.size for a in imgs)
t0 = time.perf_counter(); plan = P.ResizePlan(imgs, 64, TW); t_plan = time.perf_counter() - t0
dev = plan.host.cuda()
out = torch.empty(B, 1, 64, TW, dtype=torch.uint8, device="cuda")
ws = torch.empty(plan.workspace_bytes, dtype=torch.uint8, device="cuda")
import ctypes as C
N = pkg._native
lib = N.lib()
def launch():
    N.check(lib.i2l_resize_pad_u8(N.ptr(dev), C.c_void_p(plan.plan_ptr), C.c_void_p(dev.data_ptr() + plan.pixel_bytes),
                                  N.ptr(out), N.ptr(ws), ws.numel(), N.stream_ptr()), "resize")
for _ in range(3): launch()
torch.cuda.synchronize()
lib.i2l_prof_reset(); lib.i2l_prof_enable(1)
n = 20
for _ in range(n): launch()
torch.cuda.synchronize(); lib.i2l_prof_enable(0)
pr = N.prof_results()
tot = 0.0
for k in ("pre.resize_rows", "pre.resize_cols_pad"):
    c, ms = pr[k]; tot += ms / c
    print(f"  {k:22s} {ms / c:.4f} ms")
inter = plan.workspace_bytes
alg = src_bytes + 2 * inter + out.numel()
print(f"device resize B={B} -> 64x{TW}: {tot:.4f} ms = {B / tot * 1e3 / 1e6:.2f} M images/s; src {src_bytes / 1e6:.1f} MB, "
      f"intermediate {inter / 1e6:.1f} MB, out {out.numel() / 1e6:.1f} MB -> {alg / tot / 1e6:.0f} GB/s algorithmic; "
      f"host plan build {t_plan * 1e3:.1f} ms (incl. packing)")
# end to end from host arrays: plan + pack + H2D + kernels
for _ in range(2):                                       # both pinned staging slots exist before the clock starts
    o = P.ResizePlan(imgs, 64, TW).run("cuda")
torch.cuda.synchronize(); t0 = time.perf_counter()
for _ in range(5):
    o = P.ResizePlan(imgs, 64, TW).run("cuda")
torch.cuda.synchronize(); e2e = (time.perf_counter() - t0) / 5
print(f"end to end from host arrays (plan + pack + H2D + kernels): {e2e * 1e3:.1f} ms = {B / e2e / 1e3:.1f} k images/s")
try:
    from PIL import Image
    sub = imgs[:128]
    t0 = time.perf_counter()
    for a in sub:
        im = Image.fromarray(a, "L")
        nw = int(round(64 * (a.shape[1] / a.shape[0])))
        r = im.resize((nw, 64), Image.Resampling.LANCZOS)
        if nw < TW:
            p = Image.new("L", (TW, 64), 255); p.paste(r, (0, 0)); r = p
        elif nw > TW:
            l = (nw - TW) // 2; r = r.crop((l, 0, l + TW, 64))
        ref = np.asarray(r)
    t = time.perf_counter() - t0
    print(f"Pillow on one host core, {len(sub)} images: {t * 1e3:.1f} ms = {len(sub) / t / 1e3:.2f} k images/s")
    assert np.array_equal(ref, out[len(sub) - 1, 0].cpu().numpy())
except ImportError:
    pass
