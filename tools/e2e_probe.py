"""Host-buffer MSM timing for different streaming chunk sizes (P2B_MSM_STREAM_CHUNK hook)."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import bench
from phase2_bn254_b200 import lib
ctx = lib.Context(0)
dev = torch.device("cuda", 0)
lg = int(sys.argv[1]) if len(sys.argv) > 1 else 26
n = 1 << lg
pts = bench.make_points(torch, np, ctx, 0, n, 1, dev)
sc = bench.make_scalars(torch, n, 5, dev)
hp = torch.empty(pts.numel(), dtype=torch.uint8, pin_memory=True); hp.copy_(pts)
hs = torch.empty(sc.numel(), dtype=torch.uint8, pin_memory=True); hs.copy_(sc)
torch.cuda.synchronize()
ref = ctx.msm_dev(0, pts.data_ptr(), sc.data_ptr(), n)
t0 = time.perf_counter(); ctx.msm_dev(0, pts.data_ptr(), sc.data_ptr(), n); print("device-resident %.1f ms" % ((time.perf_counter() - t0) * 1e3))
d = torch.empty_like(pts)
torch.cuda.synchronize(); t0 = time.perf_counter(); d.copy_(hp, non_blocking=True); torch.cuda.synchronize()
print("plain H2D of the points: %.1f ms = %.1f GB/s" % ((time.perf_counter() - t0) * 1e3, pts.numel() / (time.perf_counter() - t0) / 1e9))
for chunk in [int(a) for a in sys.argv[2:]] or [22, 23, 24, 25]:
    os.environ["P2B_MSM_STREAM_CHUNK"] = str(1 << chunk)
    ts = []
    for _ in range(3):
        torch.cuda.synchronize(); t0 = time.perf_counter()
        r = ctx.msm(0, hp.numpy(), hs.numpy())
        ts.append((time.perf_counter() - t0) * 1e3)
    assert r == ref
    print("max chunk 2^%d: %s ms" % (chunk, " ".join("%.1f" % t for t in ts)), flush=True)
# where does the streamed path spend its time?  CUDA-event brackets of the library around the sort / accumulate / reduce kernels
os.environ["P2B_MSM_STREAM_CHUNK"] = str(1 << 24)
for label, fn in (("device-resident", lambda: ctx.msm_dev(0, pts.data_ptr(), sc.data_ptr(), n)), ("streamed from the host", lambda: ctx.msm(0, hp.numpy(), hs.numpy()))):
    ctx.profile(True)
    torch.cuda.synchronize(); t0 = time.perf_counter()
    fn()
    wall = (time.perf_counter() - t0) * 1e3
    prof = {name: ctx.profile_read(slot) for name, slot in (("sort", lib.PROF_MSM_SORT), ("accumulate", lib.PROF_MSM_ACCUMULATE), ("reduce", lib.PROF_MSM_REDUCE))}
    ctx.profile(False)
    print("%s: wall %.1f ms; " % (label, wall) + "; ".join("%s %.1f ms in %d kernels" % (k, v[0], v[1]) for k, v in prof.items()), flush=True)
