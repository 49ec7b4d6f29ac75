"""Multi-GPU parity check, run under torchrun on N GPUs (not a benchmark):
  * BatchedAccumulator::transform with every section split across the ranks (no collective: each rank writes its own byte
    ranges of a shared response mmap) -> the response hash must equal the single-GPU one;
  * sharded MSM (point-range shards, NCCL all-gather of the per-rank results, local sum) == single-GPU MSM.
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port 29533 tools/multi_gpu_check.py
"""
import hashlib
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
import torch.distributed as dist

import bench
from phase2_bn254_b200 import dist as pdist
from phase2_bn254_b200 import lib
from phase2_bn254_b200.powersoftau import BatchedAccumulator, CeremonyParams, PrivateKey

rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
device = torch.device("cuda", local)
dist.init_process_group("nccl", device_id=device)
ctx = lib.Context(local)
size = int(os.environ.get("POT_SIZE", "16"))
prm = CeremonyParams(size, 256)
key = PrivateKey(bench.TAU, 0x2222 * 2**190 % bench.R_MOD, 0x3333 * 2**180 % bench.R_MOD)
path_c, path_r = "/tmp/p2b_challenge.bin", "/tmp/p2b_response.bin"
if rank == 0:
    ch = np.lib.format.open_memmap(path_c, mode="w+", dtype=np.uint8, shape=(prm.accumulator_size,))
    ch[:64] = np.frombuffer(hashlib.blake2b(b"").digest(), dtype=np.uint8)
    o = 64
    g1a, g2a = np.frombuffer(bench.G1_GEN, dtype=np.uint8), np.frombuffer(bench.G2_GEN, dtype=np.uint8)
    for cnt, g in ((prm.powers_g1_length, g1a), (prm.powers_length, g2a), (prm.powers_length, g1a), (prm.powers_length, g1a), (1, g2a)):
        ch[o:o + cnt * g.size].reshape(cnt, g.size)[:] = g
        o += cnt * g.size
    ch.flush()
    rs = np.lib.format.open_memmap(path_r, mode="w+", dtype=np.uint8, shape=(prm.contribution_size,))
    rs.flush()
    del ch, rs
dist.barrier()
ch = np.load(path_c, mmap_mode="r")
rs = np.load(path_r, mmap_mode="r+")
torch.cuda.synchronize(); dist.barrier()
t0 = time.perf_counter()
pdist.sharded_transform(ctx, ch, rs, prm, key, rank, world)
rs.flush()
torch.cuda.synchronize(); dist.barrier()
dt = time.perf_counter() - t0
end = prm.contribution_size - prm.public_key_size
if rank == 0:
    sharded = hashlib.blake2b(np.load(path_r, mmap_mode="r")[64:end].tobytes()).hexdigest()
    single = np.zeros(prm.contribution_size, dtype=np.uint8)
    BatchedAccumulator.transform(ch, single, False, True, False, key, prm, ctx=ctx)
    ok = sharded == hashlib.blake2b(single[64:end].tobytes()).hexdigest()
    print("transform 2^%d over %d GPUs: %.3f s, response hash %s, equals single-GPU: %s" % (size, world, dt, sharded[:32], ok), flush=True)
    assert ok
# sharded MSM
n = 1 << 18
pts = bench.make_points(torch, np, ctx, 0, n, 1, device)        # same points on every rank
sc = bench.make_scalars(torch, n, 99, device)
lo, hi = pdist.shard_range(n, rank, world)
got = pdist.sharded_msm(ctx, 0, pts.data_ptr() + lo * 64, sc.data_ptr() + lo * 32, hi - lo, device=device, on_device=True)
whole = ctx.msm_dev(0, pts.data_ptr(), sc.data_ptr(), n)
assert got == whole, "rank %d: sharded MSM differs" % rank
if rank == 0:
    print("sharded MSM over %d GPUs equals the single-GPU result: True" % world, flush=True)
dist.barrier()
dist.destroy_process_group()
