"""Phase-2 contribute at the north-star size: MPCParameters with h = 2^k - 1 and l = 2^k points (k = 26 -> 2^27 points,
8.6 GB of wire bytes each way) on ONE GPU, host buffers in and out (the p2b_phase2_contribute call a Rust shim makes),
followed by verify_contribution of the result (two 2^k-term MSM pairs + the pairings).

    python tools/contribute_scale.py --log-m 26 [--verify]

Prints one JSON line.  Points are tau^i G generated on the GPU; h and l reuse the same vector (the cost of [delta^-1]P
does not depend on P)."""
import argparse
import hashlib
import json
import os
import struct
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--log-m", type=int, default=26)
    ap.add_argument("--reps", type=int, default=2)
    ap.add_argument("--verify", action="store_true")
    args = ap.parse_args()
    import numpy as np
    import torch
    import bench
    from phase2_bn254_b200 import lib
    from phase2_bn254_b200.phase2 import MPCParameters, verify_contribution

    device = torch.device("cuda:0")
    ctx = lib.Context(0)
    m = 1 << args.log_m
    nh, nl = m - 1, m
    t0 = time.perf_counter()
    pts = bench.make_points(torch, np, ctx, 0, m, 1, device)
    g1 = pts[:64 * 16].cpu().numpy().tobytes()
    g2 = bench.G2_GEN
    head = (g1[0:64] + g1[64:128] + g2 + g2 + bench.G1_GEN + g2 + struct.pack(">I", 2) + g1[128:256] + struct.pack(">I", nh))
    tail = (struct.pack(">I", 16) + g1[:1024] + struct.pack(">I", 16) + g1[:1024] + struct.pack(">I", 16) + g2 * 16 +
            hashlib.blake2b(b"scale").digest() + struct.pack(">I", 0))
    total = len(head) + nh * 64 + 4 + nl * 64 + len(tail)
    pin = torch.empty(total, dtype=torch.uint8, pin_memory=True)
    pout = torch.empty(total + 384, dtype=torch.uint8, pin_memory=True)
    o = 0
    pin[o:o + len(head)] = torch.frombuffer(bytearray(head), dtype=torch.uint8); o += len(head)
    pin[o:o + nh * 64].copy_(pts[:nh * 64]); o += nh * 64
    pin[o:o + 4] = torch.frombuffer(bytearray(struct.pack(">I", nl)), dtype=torch.uint8); o += 4
    pin[o:o + nl * 64].copy_(pts[:nl * 64]); o += nl * 64
    pin[o:o + len(tail)] = torch.frombuffer(bytearray(tail), dtype=torch.uint8)
    torch.cuda.synchronize()
    del pts
    setup_s = time.perf_counter() - t0
    delta = np.frombuffer(bench.be(0x2b5d1c3e7f9a0b4c6d8e0f1a2b3c4d5e6f708192a3b4c5d6e7f8091a2b3c4d5e % bench.R_MOD), dtype=np.uint8)
    s = np.frombuffer(g1[5 * 64: 6 * 64], dtype=np.uint8)
    r = np.frombuffer(lib.hash_to_g2(ctx.phase2_transcript(pin.numpy(), delta, s)), dtype=np.uint8)
    times = []
    for _ in range(args.reps):
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        _, h = ctx.phase2_contribute(pin.numpy(), delta, s, r, out=pout.numpy())
        times.append(time.perf_counter() - t0)
    out = {"workload": "phase2_contribute_2^%d" % args.log_m, "points": nh + nl, "params_bytes": total,
           "wall_s": round(min(times), 3), "wall_s_all": [round(t, 3) for t in times],
           "Mpoints_per_s": round((nh + nl) / min(times) / 1e6, 2), "setup_s": round(setup_s, 1),
           "contribution_hash": h.hex()[:32], "n_gpus": 1, "host_buffers": "pinned, in and out"}
    if args.verify:
        t0 = time.perf_counter()
        got = verify_contribution(MPCParameters(pin.numpy()), MPCParameters(pout.numpy()), ctx=ctx, rng=np.random.default_rng(1))
        out["verify_contribution_s"] = round(time.perf_counter() - t0, 3)
        out["verify_ok"] = bool(got == h)
    print(json.dumps(out))


if __name__ == "__main__":
    main()
