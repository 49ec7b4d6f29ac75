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


# ---------------------------------------------------------------------------------------------- sharded group-element FFT
_R = 21888242871839275222246405745257275088548364400416034343698204186575808495617


def _bitrev(v, bits):
    r = 0
    for _ in range(bits):
        r, v = (r << 1) | (v & 1), v >> 1
    return r


class GroupFftRank:
    """One rank's part of a radix-2 transform of d = R * L group elements, block-distributed over R = 2^k GPUs (rank r owns
    x[r L .. (r + 1) L)) -- EvaluationDomain<E, Point<G>>::{fft, ifft} (bellman/src/domain.rs:154-174,274-317) as
    prepare_phase2 uses it (powersoftau/src/bin/prepare_phase2.rs:62-105), every butterfly a 254-bit scalar multiplication.

    Decimation in frequency: the first k stages pair elements L * R / 2^(s+1) apart, i.e. on two ranks.  The pair splits the
    butterflies evenly -- the lower rank computes those of the first half of the block, the upper rank the second half
    (p2b_g*_gfft_stage: a + b and [w^pos] (a - b)) -- so every rank does L / 2 scalar multiplications per stage like in a
    single-GPU stage, and two half-block exchanges per stage move 64 / 128-byte points (a percent of the stage's compute
    time: an NCCL send / recv is all it takes).  The remaining log2(L) stages are a local transform of the block
    (p2b_g*_group_fft_scaled, scaling of the whole domain).  Rank r ends with X[R j + bitrev_k(r)], j < L.

    The exchange itself is the caller's (`exchange(stage, half, payload) -> partner's payload`): torch.distributed P2P in
    sharded_group_fft, a dictionary in the single-process simulation used by the GPU tests."""

    def __init__(self, ctx, group, block, rank, world, inverse, flags=0):
        self.ctx, self.group, self.rank, self.world, self.inverse, self.flags = ctx, group, rank, world, bool(inverse), flags
        self.k = world.bit_length() - 1
        if world < 1 or (1 << self.k) != world:
            raise ValueError("the number of ranks must be a power of two")
        self.size = _lib.enc_size(group, _lib.ENC_UNCOMPRESSED)
        self.cur = np.ascontiguousarray(_lib._host(block))
        self.L = self.cur.size // self.size
        self.log_l = self.L.bit_length() - 1
        if self.L < 2 * 1 or (1 << self.log_l) != self.L or self.cur.size != self.L * self.size:
            raise ValueError("every rank needs a power-of-two block of at least 2 points")
        self.omega = _lib.root_of_unity(self.log_l + self.k, self.inverse)

    def partner(self, s):
        return self.rank ^ (self.world >> (s + 1))

    def is_low(self, s):
        return (self.rank & (self.world >> (s + 1))) == 0

    def stage_send_inputs(self, s):
        """What the partner needs: the half of the block whose butterflies IT computes."""
        half = (self.L // 2) * self.size
        return self.cur[half:] if self.is_low(s) else self.cur[:half]

    def stage_compute(self, s, received):
        """Half of the pair's butterflies; returns the results that belong to the partner (to be sent back)."""
        bit, half = self.world >> (s + 1), (self.L // 2) * self.size
        low = self.is_low(s)
        a, b = (self.cur[:half], received) if low else (received, self.cur[half:])
        pos0 = ((self.rank & ~bit) % bit) * self.L + (0 if low else self.L // 2)
        w = pow(self.omega, 1 << s, _R).to_bytes(32, "big")
        self._sum, self._diff = self.ctx.gfft_stage(self.group, a, b, np.frombuffer(w, dtype=np.uint8), pos0,
                                                    flags=self.flags & _lib.G2_SUBGROUP)
        return self._diff if low else self._sum

    def stage_finish(self, s, received):
        """Sums stay on the lower rank, twiddled differences on the upper one."""
        self.cur = np.concatenate([self._sum, received]) if self.is_low(s) else np.concatenate([received, self._diff])
        self._sum = self._diff = None

    def finish(self):
        """The local stages; returns the rank's outputs X[R j + bitrev_k(rank)] for j < L (uncompressed wire)."""
        return self.ctx.group_fft_scaled(self.group, self.cur, self.inverse, self.log_l + self.k, flags=self.flags & _lib.G2_SUBGROUP)

    def output_offset(self):
        return _bitrev(self.rank, self.k)


def _p2p_exchange(payload, partner, device=None, group=None):
    """Swap equal-sized byte arrays with `partner` (torch.distributed P2P: NCCL on device tensors, gloo on CPU tensors)."""
    import torch
    import torch.distributed as dist
    t = torch.from_numpy(np.ascontiguousarray(payload))
    if device is not None:
        t = t.to(device)
    r = torch.empty_like(t)
    for req in dist.batch_isend_irecv([dist.P2POp(dist.isend, t, partner, group), dist.P2POp(dist.irecv, r, partner, group)]):
        req.wait()
    if device is not None:
        torch.cuda.synchronize(device)
    return r.cpu().numpy()


def sharded_group_fft(ctx, group_id, block, rank, world, inverse=False, device=None, group=None, flags=0):
    """Group FFT / iFFT of world * L points; `block` = this rank's L points x[rank L .. (rank + 1) L) (uncompressed wire).
    Returns (outputs, offset): outputs[j] = X[world * j + offset]."""
    node = GroupFftRank(ctx, group_id, block, rank, world, inverse, flags)
    for s in range(node.k):
        got = _p2p_exchange(node.stage_send_inputs(s), node.partner(s), device, group)
        back = _p2p_exchange(node.stage_compute(s, got), node.partner(s), device, group)
        node.stage_finish(s, back)
    return node.finish(), node.output_offset()


def simulate_sharded_group_fft(ctx, group_id, points, world, inverse=False, flags=0):
    """All ranks of sharded_group_fft in ONE process on one GPU (the exchanges are array hand-overs): the natural-order
    transform of `points`.  For tests of the stage arithmetic without several GPUs."""
    pts = _lib._host(points)
    size = _lib.enc_size(group_id, _lib.ENC_UNCOMPRESSED)
    d = pts.size // size
    L = d // world
    nodes = [GroupFftRank(ctx, group_id, pts[r * L * size: (r + 1) * L * size], r, world, inverse, flags) for r in range(world)]
    for s in range(nodes[0].k):
        sent = [n.stage_send_inputs(s) for n in nodes]
        back = [n.stage_compute(s, sent[n.partner(s)]) for n in nodes]
        for n in nodes:
            n.stage_finish(s, back[n.partner(s)])
    out = np.empty((L, world, size), dtype=np.uint8)
    for n in nodes:
        out[:, n.output_offset(), :] = n.finish().reshape(L, size)
    return out.reshape(-1)


def sharded_prepare_phase2(ctx, accumulator_map, parameters, m, out_map, rank, world, input_is_compressed=True,
                           check_input_for_correctness=True, device=None, group=None, g2_in_subgroup=False):
    """One iteration of powersoftau/src/bin/prepare_phase2.rs:62-241 on `world` GPUs: every rank transforms its block of the
    four vectors (sharded group iFFT above) and of the H query and writes its points of the file image phase1radix2m{m}
    into out_map (shared storage; disjoint bytes, no collective besides the transforms' exchanges).  d = 2^m >= 2 * world."""
    from .powersoftau import BatchedAccumulator, _sections
    d = 1 << m
    L = d // world
    if L < 2 or L * world != d:
        raise ValueError("need 2^m >= 2 * world")
    if len(out_map) < 192 + 384 * d:
        raise ValueError("output map too small for phase1radix2m%d" % m)
    sec = _sections(parameters, input_is_compressed)
    enc = _lib.ENC_COMPRESSED if input_is_compressed else _lib.ENC_UNCOMPRESSED
    dflags = (_lib.CHECK_INPUT if check_input_for_correctness else 0) | _lib.REJECT_INFINITY
    flags = _lib.G2_SUBGROUP if g2_in_subgroup else 0
    amap = accumulator_map

    def read(name, start, count):
        off, total, size, grp = sec[name]
        try:
            return ctx.recode(grp, np.asarray(amap[off + start * size: off + (start + count) * size]), enc, _lib.ENC_UNCOMPRESSED, dflags)
        except _lib.P2BError as e:
            BatchedAccumulator._raise(e)

    if rank == 0:                                   # header: alpha_g1, beta_g1, beta_g2
        out_map[0:64] = read("alpha_g1", 0, 1)
        out_map[64:128] = read("beta_g1", 0, 1)
        out_map[128:256] = read("beta_g2", 0, 1)
    o = 256
    for name in ("tau_g1", "tau_g2", "alpha_g1", "beta_g1"):
        grp = sec[name][3]
        size = 128 if grp else 64
        res, off = sharded_group_fft(ctx, grp, read(name, rank * L, L), rank, world, True, device, group, flags)
        out_map[o: o + d * size].reshape(L, world, size)[:, off, :] = res.reshape(L, size)
        o += d * size
    # H query: tau^(i + d) G - tau^i G for i < d - 1 (prepare_phase2.rs:132-148): plain differences on the rank's index range
    lo = rank * L
    cnt = min(L, d - 1 - lo)                        # the last rank has one pair less (and tau_g1 ends at index 2 d - 2)
    hi_pts, lo_pts = read("tau_g1", lo + d, cnt), read("tau_g1", lo, cnt)
    if cnt < L:                                     # pad to the stage's power-of-two length with a valid point; result unused
        hi_pts = np.concatenate([hi_pts, np.tile(hi_pts[:64], L - cnt)])
        lo_pts = np.concatenate([lo_pts, np.tile(lo_pts[:64], L - cnt)])
    _, h = ctx.gfft_stage(0, hi_pts, lo_pts, None, want_sum=False)
    out_map[o + lo * 64: o + (lo + cnt) * 64] = h[: cnt * 64]


def sharded_contribute(ctx, params_map, out_map, delta, s_g1, r_g2, rank, world):
    """MPCParameters::contribute with H and L split across `world` ranks (no collective: rank r rewrites only its ranges of
    out_map, shared storage such as the mapped output file; rank 0 also writes the header, the unchanged vectors and the new
    public key).  Every rank returns the same 64-byte contribution hash."""
    d = np.frombuffer(int(delta).to_bytes(32, "big"), dtype=np.uint8) if isinstance(delta, int) else delta
    return ctx.phase2_contribute(params_map, d, s_g1, r_g2, out=out_map, shard_index=rank, shard_count=world)[1]


__all__ = ["shard_range", "sharded_contribute", "GroupFftRank", "sharded_group_fft", "simulate_sharded_group_fft", "sharded_prepare_phase2", "bind_to_gpu_numa", "all_gather_bytes", "sharded_msm", "sharded_merge_pairs",
           "sharded_verify_contribution", "sharded_transform", "_lib"]
