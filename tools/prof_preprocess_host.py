import os, sys, time
sys.path.insert(0, ".")
import numpy as np, torch, i2l_import
pkg = i2l_import.load(); P = pkg.preprocess
rng = np.random.default_rng(0)
imgs = []
for i in range(1024):
    h = int(rng.integers(32, 128)); w = int(h * rng.uniform(2.0, 14.0))
    imgs.append(rng.integers(0, 256, size=(h, w), dtype=np.uint8))
for _ in range(3): P.ResizePlan(imgs, 64, 800).run("cuda")
torch.cuda.synchronize()
def t(label, fn):
    torch.cuda.synchronize(); t0 = time.perf_counter(); r = fn(); torch.cuda.synchronize(); print(f"{label}: {(time.perf_counter()-t0)*1e3:.2f} ms"); return r
pl = t("ResizePlan.__init__", lambda: P.ResizePlan(imgs, 64, 800))
t("run", lambda: pl.run("cuda"))
import ctypes as C
N = pkg._native; lib = N.lib()
arrs = [P._as_u8_array(a) for a in imgs]
t("_as_u8_array x1024", lambda: [P._as_u8_array(a) for a in imgs])
descs = (N.ImageDesc * 1024)(); off = 0
def mk():
    off = 0
    for i, a in enumerate(arrs):
        descs[i].src_offset, descs[i].height, descs[i].width = off, a.shape[0], a.shape[1]; off += (a.size + 15) // 16 * 16
t("descs loop", mk)
args = (descs, 1024, 1, 0, 64, 800, 0, 0)
nb = t("plan_bytes", lambda: lib.i2l_resize_plan_bytes(*args))
buf = torch.empty(nb, dtype=torch.uint8, pin_memory=True)
t("plan_build (pinned dst)", lambda: lib.i2l_resize_plan_build(*args, C.c_void_p(buf.data_ptr()), nb))
buf2 = torch.empty(nb, dtype=torch.uint8)
t("plan_build (pageable dst)", lambda: lib.i2l_resize_plan_build(*args, C.c_void_p(buf2.data_ptr()), nb))
t("plan_build again (pageable)", lambda: lib.i2l_resize_plan_build(*args, C.c_void_p(buf2.data_ptr()), nb))
ptrs = (C.c_void_p * 1024)(*[a.ctypes.data for a in arrs])
t("ptr array", lambda: (C.c_void_p * 1024)(*[a.ctypes.data for a in arrs]))
px = torch.empty(60_000_000, dtype=torch.uint8, pin_memory=True)
t("pack_images (pinned dst)", lambda: lib.i2l_pack_images(ptrs, descs, 1024, 1, C.c_void_p(px.data_ptr())))
host = torch.empty(81_000_000, dtype=torch.uint8, pin_memory=True)
t("H2D 81 MB", lambda: host.to("cuda", non_blocking=True))
t("torch.empty out+ws", lambda: (torch.empty(1024, 1, 64, 800, dtype=torch.uint8, device="cuda"), torch.empty(42_000_000, dtype=torch.uint8, device="cuda")))
