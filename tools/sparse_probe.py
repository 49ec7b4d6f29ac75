"""p2b_g1_sparse_mul timing: 2^20 rows x 3 entries over 2^20 bases, all-random coefficients and 90 % unit coefficients."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import bench
from phase2_bn254_b200 import lib
ctx = lib.Context(0)
nv = 1 << 20
pts = bench.make_points(torch, np, ctx, 0, nv, 1, torch.device("cuda", 0)).cpu().numpy()
rs = np.random.default_rng(11)
offs = np.arange(0, 3 * nv + 1, 3, dtype=np.uint64)
cols = rs.integers(0, nv, size=3 * nv, dtype=np.uint32)
rnd = np.frombuffer(bytearray(rs.bytes(96 * nv)), dtype=np.uint8).reshape(-1, 32).copy()
rnd[:, 0] &= 0x1f
one = np.zeros(32, dtype=np.uint8); one[31] = 1
mixed = np.tile(one, (3 * nv, 1))
gen = rs.random(3 * nv) < 0.1
mixed[gen] = rnd[gen]
for label, cf in (("all-random coefficients", rnd), ("90% unit coefficients", mixed)):
    for rep in range(2):
        t0 = time.perf_counter(); ctx.sparse_mul(0, pts, offs, cols, cf.reshape(-1)); dt = time.perf_counter() - t0
    print("%s: %.3f s = %.1f M entries/s" % (label, dt, 3 * nv / dt / 1e6), flush=True)
