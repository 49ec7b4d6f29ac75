"""G2 batch_exp at 2^17 .. 2^21 points: default (probed), P2B_G2_EXACT, P2B_G2_SUBGROUP; probe share via the profile slots."""
import json
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402
from phase2_bn254_b200 import lib  # noqa: E402

ctx = lib.Context(0)
dev = torch.device("cuda", 0)
stream = torch.cuda.ExternalStream(ctx.stream, device=dev)
k = np.frombuffer(bench.be(0x2b5d1c3e7f9a0b4c6d8e0f1a2b3c4d5e6f708192a3b4c5d6e7f8091a2b3c4d5e % bench.R_MOD), dtype=np.uint8)
tau = np.frombuffer(bench.be(bench.TAU), dtype=np.uint8)
res = {}
mmax = 1 << 21
p2 = bench.make_points(torch, np, ctx, 1, mmax, 1, dev)
o2 = torch.empty(mmax * 128, dtype=torch.uint8, device=dev)
ref = None
for lg in (17, 18, 20, 21):
    m = 1 << lg
    row = {}
    for name, fl in (("default_probed", 0), ("exact", lib.G2_EXACT), ("subgroup_flag", lib.G2_SUBGROUP)):
        ctx.batch_mul_dev(1, p2.data_ptr(), o2.data_ptr(), m, k, flags=fl); ctx.sync()
        out = o2[: m * 128].clone()
        t, _ = bench.timed(torch, stream, lambda: (ctx.batch_mul_dev(1, p2.data_ptr(), o2.data_ptr(), m, k, flags=fl), ctx.sync()), 3)
        row[name] = {"ms": round(t, 3), "Mmul_per_s": round(m / t / 1e3, 2)}
        if name == "default_probed":
            ref = out
            row["probe_verdict"] = ctx.g2_probe_stats()[1]
        else:
            row[name]["same_bytes_as_default"] = bool(torch.equal(ref, out))
    t, _ = bench.timed(torch, stream, lambda: (ctx.batch_mul_powers_dev(1, p2.data_ptr(), o2.data_ptr(), m, tau, None, 7), ctx.sync()), 2)
    row["default_probed_tau_powers"] = {"ms": round(t, 3), "Mmul_per_s": round(m / t / 1e3, 2)}
    res["g2_batch_exp_2^%d" % lg] = row
print(json.dumps(res))
