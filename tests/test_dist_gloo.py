"""CPU, world_size 2 over gloo: the N > 1 host logic -- range sharding, the all-gather of per-rank MSM results and
their combination -- with the oracle standing in for the per-rank GPU compute (checker-only use)."""
import os
import sys

import pytest
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def OracleCtx(oc):
    from util import OracleCtx as C
    return C(oc, threads=2)


def _worker(rank, world, port, q):
    sys.path.insert(0, ROOT)
    sys.path.insert(0, os.path.join(ROOT, "oracle"))
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    import torch.distributed as dist
    import oracle as oc
    from phase2_bn254_b200 import dist as pdist
    from util import random_points, random_scalars
    dist.init_process_group("gloo", init_method="tcp://127.0.0.1:%d" % port, rank=rank, world_size=world)
    try:
        res = {}
        for group, n in ((0, 301), (1, 77)):
            size = 128 if group else 64
            pts, sc = random_points(oc, group, n, seed=5, threads=2), random_scalars(n, seed=6)   # same on all ranks
            lo, hi = pdist.shard_range(n, rank, world)
            got = pdist.sharded_msm(OracleCtx(oc), group, pts[lo * size:hi * size], sc[lo * 32:hi * 32], hi - lo)
            res[group] = (got, oc.msm(group, pts, sc, threads=2))
        gathered = pdist.all_gather_bytes(bytes([rank]) * 64)
        q.put((rank, res, gathered))
    finally:
        dist.destroy_process_group()


def test_sharded_msm_world2():
    world, port = 2, 29500 + os.getpid() % 500
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, world, port, q)) for r in range(world)]
    for p in procs:
        p.start()
    out = [q.get(timeout=120) for _ in range(world)]
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    for rank, res, gathered in out:
        for group in (0, 1):
            got, exp = res[group]
            assert got == exp, "rank %d group %d" % (rank, group)
        assert gathered == bytes([0]) * 64 + bytes([1]) * 64


def _verify_worker(rank, world, port, q):
    sys.path.insert(0, ROOT)
    sys.path.insert(0, os.path.join(ROOT, "oracle"))
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    import numpy as np
    import torch.distributed as dist
    import oracle as oc
    from phase2_bn254_b200 import dist as pdist, lib
    from phase2_bn254_b200.phase2 import MPCParameters, VerificationError
    from test_verify_host import contribution_by_hand
    dist.init_process_group("gloo", init_method="tcp://127.0.0.1:%d" % port, rank=rank, world_size=world)
    try:
        lib.load()
        before, after, expected, lay = contribution_by_hand(lib, 7)          # h = 6, l = 7 points: uneven shards
        ctx = OracleCtx(oc)
        got = pdist.sharded_verify_contribution(ctx, before, MPCParameters(after), rank, world, rng=np.random.default_rng(10 + rank))
        off = lay["l"][0] + 64 * 6                                           # an element of the LAST rank's slice left as it was
        bad = bytearray(after)
        bad[off:off + 64] = before.data[off:off + 64].tobytes()
        try:
            pdist.sharded_verify_contribution(ctx, before, MPCParameters(bytes(bad)), rank, world, rng=np.random.default_rng(20 + rank))
            rejected = False
        except VerificationError:
            rejected = True
        # a point that does not decode in ONE rank's shard: that rank must still join the all-gather (no deadlock) and
        # every rank must raise (same verdict everywhere)
        hoff, hn = lay["h"][0], lay["h"][1]
        v1 = np.array(before.data[hoff: hoff + 64 * hn])
        v2 = np.array(np.frombuffer(after, dtype=np.uint8)[hoff: hoff + 64 * hn])
        v2[64 * (hn - 1)] |= 0x80                                            # flag bit in an uncompressed point: last rank's slice
        try:
            pdist.sharded_merge_pairs(ctx, v1, v2, rank, world, rng=np.random.default_rng(30 + rank))
            raised = False
        except lib.P2BError:
            raised = True
        q.put((rank, got == expected, rejected and raised))
    finally:
        dist.destroy_process_group()


def test_sharded_verify_contribution_world2():
    """The verifier's H / L random linear combinations sharded over two ranks (own random coefficients per rank, all-gather
    of the 2 x 64-byte partials): both ranks accept the hand-made contribution and both reject a tampered one."""
    world, port = 2, 30100 + os.getpid() % 500
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_verify_worker, args=(r, world, port, q)) for r in range(world)]
    for p in procs:
        p.start()
    out = [q.get(timeout=180) for _ in range(world)]
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    assert sorted(out) == [(0, True, True), (1, True, True)]


def test_shard_range_partition():
    sys.path.insert(0, ROOT)
    from phase2_bn254_b200.dist import shard_range
    for count in (0, 1, 7, 8, 2047, 1 << 20):
        for shards in (1, 2, 3, 8):
            edges = [shard_range(count, i, shards) for i in range(shards)]
            assert edges[0][0] == 0 and edges[-1][1] == count
            assert all(a[1] == b[0] for a, b in zip(edges, edges[1:]))
            sizes = [hi - lo for lo, hi in edges]
            assert max(sizes) - min(sizes) <= 1


def _gfft_worker(rank, world, port, q):
    sys.path.insert(0, ROOT)
    sys.path.insert(0, os.path.join(ROOT, "oracle"))
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    import numpy as np
    import torch.distributed as dist
    import oracle as oc
    from phase2_bn254_b200 import dist as pdist
    from util import random_points
    dist.init_process_group("gloo", init_method="tcp://127.0.0.1:%d" % port, rank=rank, world_size=world)
    try:
        ctx = OracleCtx(oc)
        res = {}
        for group, d in ((0, 16), (1, 8)):
            size = 128 if group else 64
            L = d // world
            x = random_points(oc, group, d, seed=77 + group, threads=2)                   # the same vector on every rank
            for inverse in (False, True):
                outs, off = pdist.sharded_group_fft(ctx, group, np.frombuffer(x[rank * L * size: (rank + 1) * L * size], dtype=np.uint8),
                                                    rank, world, inverse)
                whole = ctx.group_fft_scaled(group, x, inverse, d.bit_length() - 1).reshape(L, world, size)   # by definition
                res[(group, inverse)] = bool(np.array_equal(outs.reshape(L, size), whole[:, off, :]))
        q.put((rank, res))
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("world", [2, 4])
def test_sharded_group_fft_gloo(world):
    """The multi-GPU group FFT's rank logic over real torch.distributed P2P exchanges (gloo), the oracle standing in for the
    per-rank GPU kernels: every rank's outputs equal the transform by definition at X[world j + bitrev(rank)]."""
    port = 30700 + os.getpid() % 500 + world
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_gfft_worker, args=(r, world, port, q)) for r in range(world)]
    for p in procs:
        p.start()
    out = [q.get(timeout=240) for _ in range(world)]
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    for rank, res in out:
        assert all(res.values()), (rank, res)
