"""One-process-per-GPU plumbing for the parts of the hot path that shard.

  batch_exp (transform / contribute)  contiguous index ranges per rank, disjoint output bytes, NO collective
  Pippenger MSM                       shard by point range; every rank reduces its shard to ONE affine point on its
                                      GPU, the ranks all-gather those 64 / 128-byte results (NCCL has no reduce op
                                      for elliptic-curve addition) and each adds the `world` points locally
  Fr FFT                              does not shard at ceremony sizes: replicas only

The reference has no distributed layer (single process, crossbeam threads over `len / num_cpus` chunks,
powersoftau/src/batched_accumulator.rs:1137-1162, bellman/src/multicore.rs:55-71); `shard_range` is the same static
range split applied to ranks instead of threads.
"""
import numpy as np

from . import lib as _lib


def shard_range(count, index, shards):
    """[lo, hi) of shard `index`: contiguous, sizes differ by at most one, identical to the C ABI's split
    (csrc/api.cu shard_range) used by p2b_pot_transform."""
    per, rem = divmod(count, shards)
    lo = per * index + min(index, rem)
    return lo, lo + per + (1 if index < rem else 0)


def bind_to_gpu_numa(local_rank):
    """Restrict this process to the CPU cores next to its GPU (the "CPU Affinity" column of `nvidia-smi topo -m`) so that
    the pinned host buffers it allocates afterwards are first-touched on that NUMA node and H2D copies do not cross the
    socket interconnect.  Returns the core list, or None when the topology cannot be read (then nothing is changed)."""
    import os
    import subprocess
    try:
        out = subprocess.run(["nvidia-smi", "topo", "-m"], capture_output=True, text=True, timeout=20).stdout
    except (OSError, subprocess.SubprocessError):
        return None
    if "CPU Affinity" not in out:
        return None
    for line in out.splitlines():
        fields = line.split()
        if not fields or fields[0] != "GPU%d" % local_rank:
            continue
        for f in fields[1:]:                    # link types (X, NV18, SYS, ...) never start with a digit; the core list does
            if f[0].isdigit() and all(ch.isdigit() or ch in "-," for ch in f):
                cores = set()
                for part in f.split(","):
                    a, _, b = part.partition("-")
                    cores.update(range(int(a), int(b or a) + 1))
                allowed = cores & os.sched_getaffinity(0)
                if allowed:
                    os.sched_setaffinity(0, allowed)
                    return sorted(allowed)
                return None
        return None
    return None


def all_gather_bytes(payload, group=None, device=None):
    """All-gather of one fixed-size byte string per rank over torch.distributed (NCCL on GPUs, gloo on CPU).
    Returns the concatenation in rank order."""
    import torch
    import torch.distributed as dist
    world = dist.get_world_size(group)
    t = torch.frombuffer(bytearray(payload), dtype=torch.uint8)
    if device is not None:
        t = t.to(device)
    out = torch.empty(world * t.numel(), dtype=torch.uint8, device=t.device)
    dist.all_gather_into_tensor(out, t, group=group)
    return out.cpu().numpy().tobytes()


def _gather_with_status(payload, error, group=None, device=None):
    """All-gather of `payload` prefixed with a status byte.  A rank whose local GPU call failed still takes part in the
    collective (otherwise the other ranks would wait in it for ever) and flags the failure; afterwards EVERY rank raises,
    so the ranks also agree on the verdict.  Returns the payloads (status bytes stripped) as one (world, len) array."""
    size = len(payload)
    parts = np.frombuffer(all_gather_bytes(bytes([1 if error is None else 0]) + bytes(payload), group=group, device=device),
                          dtype=np.uint8).reshape(-1, 1 + size)
    bad = np.nonzero(parts[:, 0] == 0)[0]
    if bad.size:
        if error is not None:
            raise error
        raise _lib.P2BError(_lib.EDECODE, "rank %d failed on its shard" % int(bad[0]))
    return parts[:, 1:]


def sharded_msm(ctx, group_id, local_points, local_scalars, n_local, group=None, device=None, on_device=False):
    """sum over ALL ranks' terms of scalar * point.  Each rank passes only its own shard (host buffers, or device
    pointers with on_device=True); every rank returns the same uncompressed wire point."""
    size = _lib.enc_size(group_id, _lib.ENC_UNCOMPRESSED)
    part, err = bytes(size), None
    try:
        if on_device:
            part = ctx.msm_dev(group_id, local_points, local_scalars, n_local)
        else:
            pts, sc = _lib._host(local_points), _lib._host(local_scalars)
            if sc.size != 32 * n_local or pts.size != size * n_local:
                raise ValueError("sharded_msm: shard of %d terms needs %d point and %d scalar bytes" % (n_local, size * n_local,
                                                                                                    32 * n_local))
            part = ctx.msm(group_id, pts, sc)
    except (_lib.P2BError, ValueError) as e:
        err = e
    parts = _gather_with_status(part, err, group=group, device=device)
    return ctx.sum_points(group_id, np.ascontiguousarray(parts).reshape(-1))


def sharded_merge_pairs(ctx, v1, v2, rank, world, rng=None, scalar_bits=253, group=None, device=None, flags=0):
    """merge_pairs (phase2/src/utils.rs:59-105) with the index range split across ranks: every rank combines its slice of
    the two G1 vectors with its own random coefficients (one pair-MSM pass on its GPU); the ranks all-gather the 2 x 64-byte partial
    results and each adds them up.  The random coefficients need not be shared: the check is a random linear combination
    per element either way.  Every rank returns the same (s, sx); a decode failure on one rank's shard raises on all."""
    from .powersoftau import _random_scalars, system_rng
    rng = rng or system_rng()
    v1, v2 = _lib._host(v1), _lib._host(v2)
    n = v1.size // 64
    if n != v2.size // 64:
        raise ValueError("merge_pairs: length mismatch")
    lo, hi = shard_range(n, rank, world)
    zero = bytes([0x40]) + bytes(63)
    part, err = zero + zero, None
    if hi > lo:
        _random_scalars(system_rng(), 0, scalar_bits)
        try:
            a, b = ctx.msm_pair(0, v1[lo * 64: hi * 64], v2[lo * 64: hi * 64], None, bytes(rng.bytes(32)), scalar_bits, flags=flags)
            part = a + b
        except _lib.P2BError as e:
            err = e
    parts = _gather_with_status(part, err, group=group, device=device).reshape(-1, 2, 64)
    return (bytes(ctx.sum_points(0, np.ascontiguousarray(parts[:, 0]).reshape(-1))),
            bytes(ctx.sum_points(0, np.ascontiguousarray(parts[:, 1]).reshape(-1))))


def sharded_verify_contribution(ctx, before, after, rank, world, rng=None, scalar_bits=253, group=None, device=None):
    """phase2 verify_contribution with the H / L random linear combinations sharded over the ranks (the cheap checks --
    transcript, signatures of knowledge, delta ratios -- run on every rank).  Every rank returns the contribution hash or
    raises VerificationError."""
    from .phase2 import verify_contribution
    return verify_contribution(before, after, ctx=ctx, rng=rng, scalar_bits=scalar_bits,
                               merge=lambda a, b: sharded_merge_pairs(ctx, a, b, rank, world, rng, scalar_bits, group, device))


def sharded_transform(ctx, input_map, output_map, parameters, key, rank, world, input_is_compressed=False,
                      compress_the_output=True, check_input_for_correctness=False):
    """BatchedAccumulator::transform with every section split across `world` ranks (no collective: rank r writes
    only its own byte ranges of output_map, which must be shared storage such as the response mmap)."""
    from .powersoftau import BatchedAccumulator
    BatchedAccumulator.transform(input_map, output_map, input_is_compressed, compress_the_output,
                                 check_input_for_correctness, key, parameters, ctx=ctx, shard_index=rank,
                                 shard_count=world)


def sharded_contribute(ctx, params_map, out_map, delta, s_g1, r_g2, rank, world):
    """MPCParameters::contribute with H and L split across `world` ranks (no collective: rank r rewrites only its ranges of
    out_map, shared storage such as the mapped output file; rank 0 also writes the header, the unchanged vectors and the new
    public key).  Every rank returns the same 64-byte contribution hash."""
    d = np.frombuffer(int(delta).to_bytes(32, "big"), dtype=np.uint8) if isinstance(delta, int) else delta
    return ctx.phase2_contribute(params_map, d, s_g1, r_g2, out=out_map, shard_index=rank, shard_count=world)[1]


__all__ = ["shard_range", "sharded_contribute", "bind_to_gpu_numa", "all_gather_bytes", "sharded_msm", "sharded_merge_pairs",
           "sharded_verify_contribution", "sharded_transform", "_lib"]
