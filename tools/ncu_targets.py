"""Runs each kernel family once at a moderate size, for `ncu -k regex:...` captures (not a benchmark)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import bench
from phase2_bn254_b200 import lib

what = sys.argv[1] if len(sys.argv) > 1 else "all"
ctx = lib.Context(0)
dev = torch.device("cuda", 0)
k = np.frombuffer(bench.be(0x2b5d1c3e7f9a0b4c6d8e0f1a2b3c4d5e6f708192a3b4c5d6e7f8091a2b3c4d5e % bench.R_MOD), dtype=np.uint8)
if what in ("all", "g1"):
    n = 1 << 20
    p = bench.make_points(torch, np, ctx, 0, n, 1, dev)
    o = torch.empty(n * 32, dtype=torch.uint8, device=dev)
    ctx.batch_mul_powers_dev(0, p.data_ptr(), o.data_ptr(), n, k, k, 7, 0, 1); ctx.sync()      # transform shape: uncompressed in, compressed out
if what in ("all", "g2"):
    n = 1 << 18
    p = bench.make_points(torch, np, ctx, 1, n, 1, dev)
    o = torch.empty(n * 64, dtype=torch.uint8, device=dev)
    ctx.batch_mul_powers_dev(1, p.data_ptr(), o.data_ptr(), n, k, None, 7, 0, 1); ctx.sync()
if what in ("all", "fft"):
    lf = 24
    x = bench.make_scalars(torch, 1 << lf, 1, dev)
    ctx.fr_fft_dev(x.data_ptr(), lf, False, False); ctx.sync()
if what in ("all", "msm"):
    lg = int(os.environ.get("MSM_LOG", "22"))
    n = 1 << lg
    p = bench.make_points(torch, np, ctx, 0, n, 1, dev)
    s = bench.make_scalars(torch, n, 2, dev)
    print(ctx.msm_dev(0, p.data_ptr(), s.data_ptr(), n)[:8].hex())
if what == "msm_g2":
    lg = int(os.environ.get("MSM_LOG", "22"))
    n = 1 << lg
    p = bench.make_points(torch, np, ctx, 1, n, 1, dev)
    s = bench.make_scalars(torch, n, 2, dev)
    print(ctx.msm_dev(1, p.data_ptr(), s.data_ptr(), n)[:8].hex())
if what == "g2probe":                                   # one probed G2 batch_exp at 2^20 points (launch list of the probe)
    n = 1 << 20
    p = bench.make_points(torch, np, ctx, 1, n, 1, dev)
    o = torch.empty(n * 128, dtype=torch.uint8, device=dev)
    torch.cuda.nvtx.range_push("probed_batch")
    ctx.batch_mul_dev(1, p.data_ptr(), o.data_ptr(), n, k); ctx.sync()
    torch.cuda.nvtx.range_pop()
    print(ctx.g2_probe_stats())
print("done", what)
