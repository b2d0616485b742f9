"""Where does an e2e step go?  H2D bandwidth (torch pinned copy), orbfe_extract_batch alone, matcher host-view call alone."""
import sys, time, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import bench
import orb_slam_b200 as fe
from orb_slam_b200 import matching as M
B = int(sys.argv[1]) if len(sys.argv) > 1 else 64
W, H, NF = bench.W, bench.H, bench.NFEAT
frames, shifts = bench.make_stream(B, 11)
h = torch.from_numpy(frames).pin_memory()
d = torch.empty_like(h, device="cuda")
for _ in range(3): d.copy_(h, non_blocking=True)
torch.cuda.synchronize()
t = time.perf_counter()
for _ in range(10): d.copy_(h, non_blocking=True)
torch.cuda.synchronize()
dt = (time.perf_counter() - t) / 10
print("H2D %.1f MB in %.3f ms = %.1f GB/s" % (h.numel() / 1e6, dt * 1e3, h.numel() / dt / 1e9))
ex = fe.ORBextractor(NF, 1.2, 8)
mt = fe.ORBmatcher(0.9, True)
hk = torch.empty((B, NF, 28), dtype=torch.uint8).pin_memory(); hd = torch.empty((B, NF, 32), dtype=torch.uint8).pin_memory()
hc = torch.empty((B,), dtype=torch.int32).pin_memory()
def call():
    ex.extract_batch_ptr(h.data_ptr(), W, H, W, W * H, B, hk.data_ptr(), hd.data_ptr(), NF, hc.data_ptr())
for _ in range(3): call()
for rep in range(3):
    t = time.perf_counter()
    for _ in range(30): call()
    print("orbfe_extract_batch (host buffers): %.3f ms / step" % ((time.perf_counter() - t) / 30 * 1e3))
kps = hk.numpy().view(fe.KP_DTYPE).reshape(B, NF); desc = hd.numpy(); cnt = hc.numpy()
Tc = [bench.tcw_for_shift(*shifts[i]) for i in range(B)]
def match():
    t0 = time.perf_counter()
    views = [M.FrameView(kps[i, :cnt[i]], desc[i, :cnt[i]], W, H, 1.2, 8) for i in range(B)]
    lasts = [views[B - 1]] + views[:-1]
    has = [np.ones(f.n, np.uint8) for f in lasts]; outl = [np.zeros(f.n, np.uint8) for f in lasts]
    world = [bench.backproject(f.kps) for f in lasts]
    t1 = time.perf_counter()
    M.search_by_projection_frames(mt, views, lasts, has, outl, world, Tc, bench.FX, bench.FY, bench.CX, bench.CY, 15.0)
    return t1 - t0, time.perf_counter() - t1
for _ in range(3): match()
a = b = 0
for _ in range(10):
    x, y = match(); a += x; b += y
print("views+world (python): %.3f ms, orbfe_search_by_projection_frames: %.3f ms" % (a * 100, b * 100))
