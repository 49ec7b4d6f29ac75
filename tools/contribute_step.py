"""The participant step of phase 1 (compute_constrained.rs:140-230) on a 2^LOG challenge: overlapped challenge hash vs the
reference's order of operations; prints one JSON line."""
import hashlib
import json
import os
import sys
import time

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from phase2_bn254_b200 import lib  # noqa: E402
from phase2_bn254_b200.powersoftau import BatchedAccumulator, CeremonyParams, contribute_challenge  # noqa: E402

size = int(sys.argv[1]) if len(sys.argv) > 1 else 20
ctx = lib.Context(0)
prm = CeremonyParams(size, 1 << 18)
ch = torch.zeros(prm.accumulator_size, dtype=torch.uint8, pin_memory=True).numpy()
BatchedAccumulator.generate_initial(ch, False, prm)
ch[:64] = np.frombuffer(hashlib.blake2b(b"").digest(), dtype=np.uint8)
res = {"size": size, "challenge_bytes": int(ch.size)}
hashes = []
for name, ov in (("warmup", True), ("overlapped", True), ("reference_order", False)):
    rs = torch.zeros(prm.contribution_size, dtype=torch.uint8, pin_memory=True).numpy()
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    _, rh, _ = contribute_challenge(ch, rs, lib.ChaChaRng([9] * 8), prm, ctx=ctx, overlap=ov)
    res[name + "_wall_s"] = round(time.perf_counter() - t0, 4)
    hashes.append(rh)
t0 = time.perf_counter()
hashlib.blake2b(ch).digest()
res["challenge_blake2b_alone_s"] = round(time.perf_counter() - t0, 4)
t0 = time.perf_counter()
hashlib.blake2b(rs).digest()
res["response_blake2b_alone_s"] = round(time.perf_counter() - t0, 4)
rs = torch.zeros(prm.contribution_size, dtype=torch.uint8, pin_memory=True).numpy()
torch.cuda.synchronize()
t0 = time.perf_counter()
BatchedAccumulator.transform(ch, rs, False, True, False, __import__("phase2_bn254_b200.powersoftau", fromlist=["PrivateKey"]).PrivateKey(5, 6, 7), prm, ctx=ctx)
res["transform_alone_s"] = round(time.perf_counter() - t0, 4)
res["same_response"] = len(set(hashes)) == 1
res["g2_probe"] = ctx.g2_probe_stats()
print(json.dumps(res))
