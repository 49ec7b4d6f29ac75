#!/usr/bin/env python
"""bench.py -- headline benchmark of the phase2-bn254 B200 compute core.

Metric (BASELINE.json): BN254 G1 Pippenger MSM throughput in Mscalar-mul/s at 2^26 terms per GPU
(`bellman/src/multiexp.rs:330-475`), plus -- as extra keys on the same JSON line -- the
phase-2 `contribute` wall time at 2^20 constraints (config 3), batch_exp throughput (the transform hot loop) and the
Fr FFT at 2^24 (config 4).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--log-n L] [--impl ours|reference] [--no-extras]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port P bench.py --gpus N ...

One "step" = one MSM over the rank's 2^L synthetic (point, scalar) pairs.  `value` is timed with the inputs already in
HBM (CUDA events on the library's stream + barrier/synchronize on both sides, max over ranks); `e2e` is the same call
through the host-buffer C ABI entry point (`p2b_g1_msm`: pinned host buffers, H2D of 96 B/term and D2H of the 64-byte
result inside the timed region).  N > 1: weak scaling, every rank owns 2^L terms, the per-rank results are all-gathered
over NCCL and summed on every rank (phase2_bn254_b200/dist.py).  `--impl reference` times the CPU restatement of the
reference algorithm (oracle/, the reference itself is Rust and cannot be built here) on a bounded sample.
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

R_MOD = 21888242871839275222246405745257275088548364400416034343698204186575808495617
G1_GEN = (1).to_bytes(32, "big") + (2).to_bytes(32, "big")
G2_GEN = b"".join(v.to_bytes(32, "big") for v in (
    11559732032986387107991004021392285783925812861821192530917403151452391805634,
    10857046999023057135944570762232829481370756359578518086990519993285655852781,
    4082367875863433681332203403145435568316851327593401208105741076214120093531,
    8495653923123431417604973247489272438418190587263600148770280649306958101930))
METRIC = "BN254 G1 MSM throughput"
UNIT = "Mscalar-mul/s"
TAU = 0x1d7a3f6c2b9e80415f6a7b8c9d0e1f2031425364758697a8b9cadbecfd0e1f21 % R_MOD


def be(v):
    return int(v).to_bytes(32, "big")


def host_cores():
    try:
        return len(os.sched_getaffinity(0))
    except AttributeError:
        return os.cpu_count() or 1


# ----------------------------------------------------------------------------------------------- CPU legs (oracle)
def bellman_windows(n):
    """bellman's window rule (bellman/src/multiexp.rs:341-345): c = 3 if n < 32 else ceil(ln n); windows while skip < 254."""
    import math
    c = 3 if n < 32 else int(math.ceil(math.log(n)))
    return c, (254 + c - 1) // c


def cpu_points_scalars(n, seed=1):
    """n G1 points and n uniform scalars for the CPU arm.  Above 2^16 terms the 2^16 distinct points tau^(i+1) G are tiled:
    Pippenger's cost per term does not depend on the point values (every term is one mixed add into a bucket chosen by the
    scalar), and generating 2^26 distinct points with the CPU restatement's batch_exp would take minutes."""
    sys.path.insert(0, os.path.join(ROOT, "oracle"))
    import numpy as np
    import oracle as oc
    distinct = min(n, 1 << 16)
    base = np.frombuffer(oc.batch_mul_powers(0, G1_GEN * distinct, be(TAU), None, 1, threads=host_cores()), dtype=np.uint8)
    pts = np.tile(base, n // distinct) if n > distinct else base
    rng = np.random.default_rng(seed)
    sc = np.empty((n, 32), dtype=np.uint8)
    step = 1 << 22
    for lo in range(0, n, step):
        sc[lo:lo + step] = rng.integers(0, 256, size=(min(step, n - lo), 32), dtype=np.uint8)
    sc[:, 0] &= 0x1f
    return pts, sc.reshape(-1)


def cpu_msm(pts, sc, cores):
    """One oracle Pippenger over host arrays: (result bytes, decode seconds, Pippenger seconds)."""
    import oracle as oc
    return oc.msm_timed(0, pts, sc, threads=cores)


def run_reference(args, rank):
    """The CPU arm: the oracle's restatement of bellman's multiexp (the reference itself is Rust and cannot be built here) on
    the SAME config as the GPU arm -- 2^log_n terms per step, bellman's own window width for that size -- on all host
    threads.  A step is one full-size MSM (about a minute at 2^26 on 16 cores), so the number of steps actually run is
    bounded by a time budget and reported; the timed figure is the Pippenger proper (the wire decode of the inputs, which
    the reference's multiexp does not do, is reported separately)."""
    if rank != 0:
        return
    log_n = args.ref_log_n if args.ref_log_n else args.log_n
    n = 1 << log_n
    cores = host_cores()
    pts, sc = cpu_points_scalars(n)
    c, nwin = bellman_windows(n)
    budget_s, t_start, times, dec = args.ref_budget_s, time.perf_counter(), [], 0.0
    while len(times) < max(1, args.steps):
        _, d, t = cpu_msm(pts, sc, cores)
        times.append(t)
        dec = d
        if time.perf_counter() - t_start + t > budget_s:
            break
    dt = sum(times) / len(times)
    rate = n / dt / 1e6
    line = {
        "impl": "reference", "metric": METRIC, "value": round(rate, 4), "unit": UNIT, "n_gpus": args.gpus,
        "steps": len(times), "warmup": 0, "ms_per_step": round(dt * 1e3, 3), "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "u64", "data": "synthetic",
        "config": {"workload": "g1_msm_2^%d" % log_n, "terms_per_gpu": n, "window_bits": c, "windows": nwin,
                   "steps_requested": args.steps, "time_budget_s": budget_s},
        "cpu_baseline": {"value": round(rate, 4), "unit": UNIT, "cores": cores, "kind": "port",
                         "sample": "%d full Pippenger MSM(s) over 2^%d terms (bellman window rule c = ceil(ln n) = %d, %d windows, "
                                   "one task per window and point chunk on all %d host threads); Pippenger proper timed, wire "
                                   "decode of the inputs (%.1f s, not part of the reference's multiexp) excluded; the reference is "
                                   "Rust and cannot be built here, this is oracle/p2b_oracle.c; 2^16 distinct points tiled"
                                   % (len(times), log_n, c, nwin, cores, dec)},
        "e2e": {"value": round(rate, 4), "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }
    print(json.dumps(line), flush=True)


# ----------------------------------------------------------------------------------------------- clocks sampler
class ClockSampler:
    FIELDS = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
              "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.samples, self.proc, self.thread = [], None, None
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(index), "--query-gpu=" + self.FIELDS,
                                          "--format=csv,noheader,nounits", "-lms", "100"], stdout=subprocess.PIPE,
                                         stderr=subprocess.DEVNULL, text=True)
            self.thread = threading.Thread(target=self._read, daemon=True)
            self.thread.start()
        except OSError:
            self.proc = None
        self.mark = 0

    def _read(self):
        for line in self.proc.stdout:
            self.samples.append((time.perf_counter(), line.strip()))

    def window(self, t0, t1):
        rows = [s for t, s in self.samples if t0 <= t <= t1] or [s for _, s in self.samples[-3:]]
        sm, mx, reasons = [], 0, set()
        for r in rows:
            f = [x.strip() for x in r.split(",")]
            try:
                sm.append(float(f[0])); mx = max(mx, float(f[1]))
            except (ValueError, IndexError):
                continue
            for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), f[3:7]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        sm.sort()
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": mx or None, "reasons": sorted(reasons),
                "samples": len(sm)}

    def stop(self):
        if self.proc:
            self.proc.terminate()


# ----------------------------------------------------------------------------------------------- GPU workload
def msm_windows(log_n):
    """Number of windows csrc/msm_impl.cuh msm_geometry() uses for 2^log_n terms (for the field-multiplication count)."""
    target = max(6, min(20, int(0.6 * log_n + 3.5)))
    for c in range(target, 1, -1):
        nwin = (255 + c - 1) // c
        if (255 + nwin - 1) // nwin == c:
            return nwin
    return 15


def make_scalars(torch, n, seed, device):
    """n x 32 BE bytes: uniform 254-bit values folded into [0, r) (values >= r get their top nibble cleared)."""
    rbytes = torch.tensor(list(be(R_MOD)), dtype=torch.int16, device=device)
    g = torch.Generator(device=device)
    g.manual_seed(seed)
    out = torch.empty((n, 32), dtype=torch.uint8, device=device)
    step = 1 << 22
    for lo in range(0, n, step):
        m = min(step, n - lo)
        s = torch.randint(0, 256, (m, 32), dtype=torch.uint8, device=device, generator=g)
        s[:, 0] &= 0x3f
        d = s.to(torch.int16) - rbytes
        nz = d != 0
        first = nz.to(torch.uint8).argmax(dim=1, keepdim=True)
        ge = (d.gather(1, first).squeeze(1) > 0) | (~nz.any(dim=1))
        s[ge, 0] &= 0x0f
        out[lo:lo + m] = s
    return out.reshape(-1)


def make_points(torch, np, ctx, group, n, start, device):
    """n points tau^(start+i) * G generated on the GPU by the library's own batch_exp (a valid powers-of-tau vector)."""
    gen = G2_GEN if group else G1_GEN
    src = torch.from_numpy(np.frombuffer(gen, dtype=np.uint8).copy()).to(device).repeat(n)
    pts = torch.empty(n * len(gen), dtype=torch.uint8, device=device)
    torch.cuda.synchronize()
    ctx.batch_mul_powers_dev(group, src.data_ptr(), pts.data_ptr(), n, np.frombuffer(be(TAU), dtype=np.uint8), None, start)
    ctx.sync()
    del src
    return pts


def pageable_msm(args, torch, np, ctx, np_pts, np_sc, m):
    """e2e MSM over m terms read from memory maps of real files (pageable: staged through the library's pinned rings by
    host threads, csrc/hostio.cu) next to the same call on page-locked buffers."""
    import tempfile
    out = {}
    d = tempfile.mkdtemp(prefix="p2b_bench_", dir="/dev/shm" if os.path.isdir("/dev/shm") and args.mmap_dir is None else args.mmap_dir)
    try:
        fp, fs = os.path.join(d, "points"), os.path.join(d, "scalars")
        np_pts[: 64 * m].tofile(fp)
        np_sc[: 32 * m].tofile(fs)
        pm, sm = np.memmap(fp, dtype=np.uint8, mode="r"), np.memmap(fs, dtype=np.uint8, mode="r")
        res = {}
        for name, a, b in (("pinned", np_pts[: 64 * m], np_sc[: 32 * m]), ("mmap", pm, sm)):
            ctx.msm(0, a, b)
            ts = []
            for _ in range(3):
                torch.cuda.synchronize()
                t0 = time.perf_counter()
                res[name] = ctx.msm(0, a, b)
                ts.append(time.perf_counter() - t0)
            out[name + "_ms"] = round(min(ts) * 1e3, 3)
        out["terms"] = m
        out["mmap_over_pinned"] = round(out["mmap_ms"] / out["pinned_ms"], 3)
        out["same_result"] = res["pinned"] == res["mmap"]
        out["files_in"] = d.rsplit("/", 1)[0]
        del pm, sm
    finally:
        import shutil
        shutil.rmtree(d, ignore_errors=True)
    return out


def scaling_extras(args, torch, np, dist, pdist, lib, ctx, rank, world, device, e2e_ms, n_headline):
    """Multi-GPU rows (BASELINE config 5, the sharded transform / contribute, the north-star contribute at 2^26 constraints) and
    the host-side ceiling that explains the end-to-end scaling.  Run by every rank; returns the record on rank 0.
    Wall-clock between barriers, max over ranks, for the host-buffer paths; shared files live in --mmap-dir (/dev/shm)."""
    import hashlib
    import shutil
    import struct
    import tempfile
    from phase2_bn254_b200.powersoftau import BatchedAccumulator, CeremonyParams, PrivateKey
    out = {}

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def max_over_ranks(v):
        if world == 1:
            return float(v)
        t = torch.tensor([v], dtype=torch.float64, device=device)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    def gather_floats(v):
        if world == 1:
            return [float(v)]
        t = torch.tensor([v], dtype=torch.float64, device=device)
        o = torch.empty(world, dtype=torch.float64, device=device)
        dist.all_gather_into_tensor(o, t)
        return [float(x) for x in o.cpu()]

    def bcast(obj):
        if world == 1:
            return obj
        box = [obj]
        dist.broadcast_object_list(box, src=0)
        return box[0]

    def wall(fn, reps=1):
        best = 1e30
        for _ in range(reps):
            barrier()
            t0 = time.perf_counter()
            fn()
            torch.cuda.synchronize()
            best = min(best, max_over_ranks(time.perf_counter() - t0))
        return best

    # -- (1) raw concurrent pinned host -> device bandwidth of this box, all ranks at once: the ceiling of every e2e number
    nb = 1 << 30
    hbuf = torch.empty(nb, dtype=torch.uint8, pin_memory=True)
    hbuf.zero_()
    dbuf = torch.empty(nb, dtype=torch.uint8, device=device)
    dbuf.copy_(hbuf, non_blocking=True)
    torch.cuda.synchronize()
    barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(4):
        dbuf.copy_(hbuf, non_blocking=True)
    e1.record()
    torch.cuda.synchronize()
    per_rank = gather_floats(4 * nb / (e0.elapsed_time(e1) * 1e-3) / 1e9)
    del hbuf, dbuf
    need = 96.0 * n_headline / (e2e_ms * 1e-3) / 1e9
    out["h2d_ceiling"] = {"pinned_GBps_per_rank_concurrent": [round(x, 1) for x in per_rank], "aggregate_GBps": round(sum(per_rank), 1),
                          "e2e_msm_GBps_per_rank_achieved": round(need, 1),
                          "note": "every rank copies 4 x 1 GiB pinned -> device at the same time (CUDA events); the e2e MSM moves 96 B per "
                                  "term per step from pinned host memory on every rank at once"}

    # -- (2) BASELINE config 5: G1 + G2 Pippenger, 2^25 terms per GPU (2^28 in total on 8 GPUs), point-range shards, all-gather of
    #        the per-rank results + local sum; device-resident
    lg5 = args.config5_log_n
    n5 = 1 << lg5
    sc5 = make_scalars(torch, n5, 0x5dbe6259 + rank, device)
    res5 = {}
    for grp, name in ((0, "g1"), (1, "g2")):
        p5 = make_points(torch, np, ctx, grp, n5, 1 + rank * n5, device)
        box = {}

        def step():
            box["r"] = pdist.sharded_msm(ctx, grp, p5.data_ptr(), sc5.data_ptr(), n5, device=device, on_device=True) if world > 1 \
                else ctx.msm_dev(grp, p5.data_ptr(), sc5.data_ptr(), n5)

        step()
        t = wall(step, 3)
        part = ctx.msm_dev(grp, p5.data_ptr(), sc5.data_ptr(), n5)
        parts = pdist.all_gather_bytes(part, device=device) if world > 1 else part
        res5[name] = {"ms": round(t * 1e3, 3), "Mscalar_mul_per_s": round(world * n5 / t / 1e6, 1),
                      "equals_sum_of_per_rank_results": bool(ctx.sum_points(grp, np.frombuffer(parts, dtype=np.uint8)) == box["r"]),
                      "result_prefix": box["r"][:16].hex()}
        # cross-check at 2^20 terms per rank: the sharded sum equals ONE GPU's MSM over all world * 2^20 terms (rank 0 regenerates them)
        m = 1 << 20
        small = pdist.sharded_msm(ctx, grp, p5.data_ptr(), sc5.data_ptr(), m, device=device, on_device=True) if world > 1 \
            else ctx.msm_dev(grp, p5.data_ptr(), sc5.data_ptr(), m)
        if world > 1:                        # every rank's first 2^20 points and scalars, gathered over NCCL
            esz = 128 if grp else 64
            allp = torch.empty(world * m * esz, dtype=torch.uint8, device=device)
            alls = torch.empty(world * m * 32, dtype=torch.uint8, device=device)
            dist.all_gather_into_tensor(allp, p5[: m * esz].contiguous())
            dist.all_gather_into_tensor(alls, sc5[: m * 32].contiguous())
            torch.cuda.synchronize()
            if rank == 0:
                res5[name]["sharded_equals_single_gpu_at_2^20_per_rank"] = bool(
                    ctx.msm_dev(grp, allp.data_ptr(), alls.data_ptr(), world * m) == small)
            del allp, alls
        del p5
        torch.cuda.empty_cache()
    del sc5
    out["config5_g1_g2_msm_2^%d_per_gpu" % lg5] = dict(res5, total_terms_per_group=world * n5,
                                                        combined_ms=round(res5["g1"]["ms"] + res5["g2"]["ms"], 3))

    # -- shared files for the sharded host-buffer rows
    base = args.mmap_dir or ("/dev/shm" if os.path.isdir("/dev/shm") else tempfile.gettempdir())
    d = bcast(tempfile.mkdtemp(prefix="p2b_scale_", dir=base) if rank == 0 else None)
    free_gb = shutil.disk_usage(d).free / 1e9
    try:
        # -- (3) BatchedAccumulator::transform at 2^22 powers, every section split across the ranks, disjoint writes into ONE
        #        response file; the bytes must equal the single-GPU response
        size = args.sharded_transform_size
        prm = CeremonyParams(size, 256)
        key = PrivateKey(TAU, 0x2222 * 2**190 % R_MOD, 0x3333 * 2**180 % R_MOD)
        cf, rf, rf1 = os.path.join(d, "challenge"), os.path.join(d, "response"), os.path.join(d, "response_n1")
        if rank == 0:
            ch = np.memmap(cf, dtype=np.uint8, mode="w+", shape=(prm.accumulator_size,))
            ch[:64] = np.frombuffer(hashlib.blake2b(b"").digest(), dtype=np.uint8)
            BatchedAccumulator.generate_initial(ch, False, prm)
            ch.flush()
            del ch
            for f in (rf, rf1):
                with open(f, "wb") as fh:
                    fh.truncate(prm.contribution_size)
        barrier()
        cm = np.memmap(cf, dtype=np.uint8, mode="r")
        rm = np.memmap(rf, dtype=np.uint8, mode="r+")
        # warm-up on a 2^10 accumulator: module load of the batch_exp kernels and the pinned staging rings of this context
        p10 = CeremonyParams(10, 256)
        w_in = np.zeros(p10.accumulator_size, dtype=np.uint8)
        BatchedAccumulator.generate_initial(w_in, False, p10)
        BatchedAccumulator.transform(w_in, np.zeros(p10.contribution_size, dtype=np.uint8), False, True, False, key, p10, ctx=ctx)
        t = wall(lambda: pdist.sharded_transform(ctx, cm, rm, prm, key, rank, world), 2)
        rm.flush()
        barrier()
        row = {"wall_s": round(t, 4), "ranks": world, "challenge_bytes": prm.accumulator_size}
        if rank == 0:
            end = prm.contribution_size - prm.public_key_size
            body = np.memmap(rf, dtype=np.uint8, mode="r")[64:end]
            row["response_body_blake2b"] = hashlib.blake2b(body).hexdigest()[:32]
            if world > 1:
                r1 = np.memmap(rf1, dtype=np.uint8, mode="r+")
                t0 = time.perf_counter()
                BatchedAccumulator.transform(cm, r1, False, True, False, key, prm, ctx=ctx)
                row["single_gpu_wall_s"] = round(time.perf_counter() - t0, 4)
                row["equals_single_gpu_response"] = bool(np.array_equal(r1[64:end], body))
                del r1
            del body
        out["sharded_transform_2^%d" % size] = row
        # -- (3b) prepare_phase2 for m = 20 from that response (compressed): four group iFFTs of 2^20 points (3 G1 + 1 G2) sharded
        #         over the ranks (rank-crossing butterfly stages exchanged over NCCL P2P, block-local transforms) + the H query
        mp2 = min(args.prepare_m, size)
        qf = os.path.join(d, "radix")
        qbytes = 192 + 384 * (1 << mp2)
        if rank == 0:
            with open(qf, "wb") as fh:
                fh.truncate(qbytes)
        barrier()
        rro = np.memmap(rf, dtype=np.uint8, mode="r")
        qm = np.memmap(qf, dtype=np.uint8, mode="r+")
        if (1 << mp2) >= 2 * world:
            # warm-up at m = 6: NCCL sets up its P2P connections on the first send / recv between a pair of ranks
            pdist.sharded_prepare_phase2(ctx, rro, prm, 6, np.zeros(192 + 384 * 64, dtype=np.uint8), rank, world, True, True,
                                         device if world > 1 else None)
            t = wall(lambda: pdist.sharded_prepare_phase2(ctx, rro, prm, mp2, qm, rank, world, True, True, device if world > 1 else None))
            qm.flush()
            barrier()
            row = {"wall_s": round(t, 4), "ranks": world, "file_bytes": qbytes,
                   "scalar_muls": 4 * (1 << (mp2 - 1)) * mp2 + 4 * (1 << mp2)}
            if rank == 0:
                img = np.memmap(qf, dtype=np.uint8, mode="r")
                row["file_blake2b"] = hashlib.blake2b(img).hexdigest()[:32]
                t0 = time.perf_counter()
                one = ctx.pot_prepare_phase2(rro, size, mp2, compressed_input=True)
                row["single_gpu_call_wall_s"] = round(time.perf_counter() - t0, 4)
                row["equals_single_gpu_file"] = bool(np.array_equal(one, img))
                del img, one
            out["sharded_prepare_phase2_m%d" % mp2] = row
        del rro, qm
        del cm, rm
        barrier()
        if rank == 0:
            os.remove(qf)
        if rank == 0:
            for f in (cf, rf, rf1):
                os.remove(f)

        # -- (4) MPCParameters::contribute with H and L split across the ranks: 2^24 constraints, and the north-star size 2^26
        #        (2^27 H / L points, 8.6 GB of parameters each way) when the shared directory has room for the files
        delta = np.frombuffer(be(0x2b5d1c3e7f9a0b4c6d8e0f1a2b3c4d5e6f708192a3b4c5d6e7f8091a2b3c4d5e % R_MOD), dtype=np.uint8)
        for lgm in (24, 26):
            m = 1 << lgm
            nh, nl = m - 1, m
            total = 64 * 2 + 128 * 2 + 64 + 128 + 4 + 128 + 4 + nh * 64 + 4 + nl * 64 + (4 + 1024) * 2 + 4 + 2048 + 64 + 4
            if lgm > args.max_contribute_log or free_gb < 3.2 * total / 1e9 + 2:
                out["sharded_contribute_2^%d" % lgm] = {"skipped": "needs %.1f GB in %s (free: %.1f GB) or above --max-contribute-log"
                                                                   % (3.2 * total / 1e9, base, free_gb)}
                continue
            pf, of, of1 = os.path.join(d, "params"), os.path.join(d, "params_out"), os.path.join(d, "params_out_n1")
            if rank == 0:
                ptsd = make_points(torch, np, ctx, 0, m, 1, device)
                g1 = ptsd[: 64 * 16].cpu().numpy().tobytes()
                head = (g1[0:64] + g1[64:128] + G2_GEN + G2_GEN + G1_GEN + G2_GEN + struct.pack(">I", 2) + g1[128:256] +
                        struct.pack(">I", nh))
                tail = (struct.pack(">I", 16) + g1[:1024] + struct.pack(">I", 16) + g1[:1024] + struct.pack(">I", 16) + G2_GEN * 16 +
                        hashlib.blake2b(b"scale").digest() + struct.pack(">I", 0))
                assert len(head) + nh * 64 + 4 + nl * 64 + len(tail) == total
                pm = np.memmap(pf, dtype=np.uint8, mode="w+", shape=(total,))
                hp = ptsd.cpu().numpy()
                o = 0
                pm[o:o + len(head)] = np.frombuffer(head, dtype=np.uint8); o += len(head)
                pm[o:o + nh * 64] = hp[: nh * 64]; o += nh * 64
                pm[o:o + 4] = np.frombuffer(struct.pack(">I", nl), dtype=np.uint8); o += 4
                pm[o:o + nl * 64] = hp[: nl * 64]; o += nl * 64
                pm[o:o + len(tail)] = np.frombuffer(tail, dtype=np.uint8)
                pm.flush()
                s_g1 = np.frombuffer(g1[5 * 64: 6 * 64], dtype=np.uint8).copy()
                r_g2 = np.frombuffer(lib.hash_to_g2(ctx.phase2_transcript(pm, delta, s_g1)), dtype=np.uint8).copy()
                del pm, hp, ptsd
                torch.cuda.empty_cache()
                for f in (of,) + ((of1,) if world > 1 else ()):
                    with open(f, "wb") as fh:
                        fh.truncate(total + 384)
                keys = (s_g1.tobytes(), r_g2.tobytes())
            else:
                keys = None
            keys = bcast(keys)
            s_g1, r_g2 = np.frombuffer(keys[0], dtype=np.uint8), np.frombuffer(keys[1], dtype=np.uint8)
            barrier()
            pm = np.memmap(pf, dtype=np.uint8, mode="r")
            om = np.memmap(of, dtype=np.uint8, mode="r+")
            box = {}

            def step():
                box["h"] = pdist.sharded_contribute(ctx, pm, om, delta, s_g1, r_g2, rank, world)

            t = wall(step, 2)
            row = {"wall_s": round(t, 4), "ranks": world, "points": nh + nl, "params_bytes": total,
                   "Mpoints_per_s": round((nh + nl) / t / 1e6, 1), "contribution_hash": box["h"].hex()[:32],
                   "buffers": "memory maps of files in %s (pageable; staged through pinned rings)" % base}
            if rank == 0 and world > 1:
                o1 = np.memmap(of1, dtype=np.uint8, mode="r+")
                t0 = time.perf_counter()
                _, h1 = ctx.phase2_contribute(pm, delta, s_g1, r_g2, out=o1)
                row["single_gpu_wall_s"] = round(time.perf_counter() - t0, 4)
                row["equals_single_gpu_file"] = bool(h1 == box["h"] and np.array_equal(o1, np.memmap(of, dtype=np.uint8, mode="r")))
                del o1
            out["sharded_contribute_2^%d" % lgm] = row
            del pm, om
            barrier()
            if rank == 0:
                for f in (pf, of, of1):
                    if os.path.exists(f):
                        os.remove(f)
    finally:
        barrier()
        if rank == 0:
            shutil.rmtree(d, ignore_errors=True)
    return out if rank == 0 else None


def timed(torch, stream, fn, reps):
    """fn() `reps` times between two CUDA events on the library's stream; returns (device ms, wall ms) per rep."""
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    e0.record(stream)
    for _ in range(reps):
        fn()
    e1.record(stream)
    torch.cuda.synchronize()
    wall = (time.perf_counter() - t0) * 1e3 / reps
    return e0.elapsed_time(e1) / reps, wall


def run_ours(args, rank, world, local_rank):
    import numpy as np
    import torch
    import torch.distributed as dist
    from phase2_bn254_b200 import dist as pdist
    from phase2_bn254_b200 import lib

    numa_cores = pdist.bind_to_gpu_numa(local_rank) if world > 1 else None    # pinned buffers land next to the GPU
    torch.cuda.set_device(local_rank)
    device = torch.device("cuda", local_rank)
    # stdout carries exactly one JSON line: anything native libraries print there (e.g. NCCL's version banner) goes to stderr
    sys.stdout.flush()
    real_stdout = os.dup(1)
    os.dup2(2, 1)
    if world > 1:
        dist.init_process_group("nccl", device_id=device)
    ctx = lib.Context(local_rank)
    stream = torch.cuda.ExternalStream(ctx.stream, device=device)
    n = 1 << args.log_n
    peaks = {}
    try:
        peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
    except (OSError, ValueError):
        pass
    hbm_peak, peak_src = (peaks["hbm_gbs"], "measured (MEASURED_PEAKS.json)") if "hbm_gbs" in peaks else (6650.0, "fallback")

    # ---- synthetic inputs, generated on the device: rank r owns terms [r n, (r+1) n)
    pts = make_points(torch, np, ctx, 0, n, 1 + rank * n, device)
    sc = make_scalars(torch, n, 0x8d313d76 + rank, device)
    torch.cuda.synchronize()

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    result = {}

    def step_dev():
        if world > 1:
            result["r"] = pdist.sharded_msm(ctx, 0, pts.data_ptr(), sc.data_ptr(), n, device=device, on_device=True)
        else:
            result["r"] = ctx.msm_dev(0, pts.data_ptr(), sc.data_ptr(), n)

    clocks = ClockSampler(local_rank) if rank == 0 else None
    for _ in range(args.warmup):
        step_dev()
    ctx.profile(True)
    launches0 = ctx.launch_count
    barrier()
    t_w0 = time.perf_counter()
    dev_ms, wall_ms = timed(torch, stream, step_dev, args.steps)
    barrier()
    t_w1 = time.perf_counter()
    launches = ctx.launch_count - launches0
    prof = {name: ctx.profile_read(slot) for name, slot in (("sort", lib.PROF_MSM_SORT), ("accumulate", lib.PROF_MSM_ACCUMULATE),
                                                            ("reduce", lib.PROF_MSM_REDUCE))}
    ctx.profile(False)
    # a multi-rank step also has NCCL work on another stream: the wall clock between the barriers is the step time there
    step_ms = wall_ms if world > 1 else dev_ms
    if world > 1:
        t = torch.tensor([step_ms], dtype=torch.float64, device=device)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        step_ms = float(t.item())
    value = world * n / (step_ms * 1e-3) / 1e6

    # ---- end to end: host (pinned) buffers through the host entry point, H2D + D2H inside the timed region
    h_pts = torch.empty(pts.numel(), dtype=torch.uint8, pin_memory=True)
    h_sc = torch.empty(sc.numel(), dtype=torch.uint8, pin_memory=True)
    h_pts.copy_(pts); h_sc.copy_(sc)
    torch.cuda.synchronize()
    np_pts, np_sc = h_pts.numpy(), h_sc.numpy()

    def step_host():
        if world > 1:
            result["e"] = pdist.sharded_msm(ctx, 0, np_pts, np_sc, n, device=device)
        else:
            result["e"] = ctx.msm(0, np_pts, np_sc)

    for _ in range(min(args.warmup, 2)):
        step_host()
    barrier()
    _, e2e_ms = timed(torch, stream, step_host, args.steps)
    barrier()
    if world > 1:
        t = torch.tensor([e2e_ms], dtype=torch.float64, device=device)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        e2e_ms = float(t.item())
    e2e_value = world * n / (e2e_ms * 1e-3) / 1e6
    assert result["e"] == result["r"], "host-buffer and device-buffer MSM disagree"
    # ---- the CPU restatement on the GPU arm's OWN inputs, full size (rank 0, one GPU): the parity check at the size the
    #      metric is quoted on, and the cpu_baseline of this line
    cpu_line = None
    if world == 1 and rank == 0 and not args.no_cpu:
        sys.path.insert(0, os.path.join(ROOT, "oracle"))
        cores = host_cores()
        cres, cdec, cdt = cpu_msm(np_pts, np_sc, cores)
        c_bits, c_win = bellman_windows(n)
        cpu_line = {"value": round(n / cdt / 1e6, 4), "unit": UNIT, "cores": cores, "kind": "port",
                    "sample": "ONE full Pippenger MSM over the GPU arm's own 2^%d points and scalars (%.1f s; bellman window rule "
                              "c = %d, %d windows; wire decode of the inputs, %.1f s, excluded), oracle/p2b_oracle.c = C restatement "
                              "of bellman multiexp on all host threads" % (args.log_n, cdt, c_bits, c_win, cdec),
                    "result_matches_gpu": cres == result["r"]}
        assert cres == result["r"], "GPU MSM at 2^%d differs from the oracle" % args.log_n
    # ---- pageable caller buffers (memory maps of real files, as the reference's binaries pass them): MSM 2^24 from an
    #      np.memmap next to the same call on the pinned copy
    pageable = None
    if world == 1 and rank == 0 and not args.no_extras:
        pageable = pageable_msm(args, torch, np, ctx, np_pts, np_sc, min(n, 1 << 24))
    del h_pts, h_sc, np_pts, np_sc

    # ---- rows that only exist with several GPUs, or whose single-GPU figure is the reference point for them (all ranks take
    #      part; rank 0 keeps the record)
    scale_rows = None
    scale_failed = False
    if not args.no_extras:
        del pts, sc
        torch.cuda.empty_cache()
        pts = sc = None
        # The headline numbers above are already measured: a failure (or a hang) in these additional rows must not lose them.
        import signal

        def _alarm(signum, frame):
            raise TimeoutError("scaling rows exceeded %d s" % args.scaling_rows_timeout)

        signal.signal(signal.SIGALRM, _alarm)
        signal.alarm(args.scaling_rows_timeout)
        try:
            scale_rows = scaling_extras(args, torch, np, dist, pdist, lib, ctx, rank, world, device, e2e_ms, n)
        except BaseException as exc:       # noqa: BLE001 -- reported in the record; the ranks may be out of step afterwards
            scale_rows = {"error": "%s: %s" % (type(exc).__name__, exc)}
            scale_failed = True
            if rank != 0:
                os._exit(0)
        finally:
            signal.alarm(0)

    if rank != 0:
        if world > 1:
            dist.barrier()
            dist.destroy_process_group()
        return

    acc_ms, acc_k = prof["accumulate"]
    acc_avg_ms = acc_ms / max(1, args.steps)       # the accumulation of one MSM (one launch per window group)
    achieved = 96.0 * n / (acc_avg_ms * 1e-3) / 1e9 if acc_avg_ms else None
    traffic = pipe_busy = mul_peak = capture = None
    try:
        prof_json = json.load(open(os.path.join(ROOT, "profiles", "traffic.json")))
        mul_peak = prof_json.get("mont_mul_peak_gmul_per_s")
        key = "k_msm_accumulate<Fq>@2^%d" % args.log_n
        capture = prof_json.get(key + ":capture")
        # the ncu figures are only printed when they were captured from the kernel sources this run was built from
        import hashlib
        h = hashlib.sha256()
        for f in (capture or {}).get("sources", []):
            h.update(open(os.path.join(ROOT, f), "rb").read())
        if capture and h.hexdigest()[:16] == capture.get("source_sha256_16"):
            traffic, pipe_busy = prof_json.get(key), prof_json.get(key + ":fmaheavy_busy")
        elif capture:
            capture = dict(capture, stale="kernel sources changed since the capture: traffic / busy_frac withheld")
    except (OSError, ValueError):
        pass
    line = {
        "metric": METRIC, "value": round(value, 3), "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": round(step_ms, 4), "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "u32", "data": "synthetic",
        "config": {"workload": "g1_msm_2^%d" % args.log_n, "terms_per_gpu": n, "points": "tau^i*G (uncompressed wire, 64 B)",
                   "scalars": "uniform mod r (32 B BE)", "arithmetic": "8 x u32 limbs, 256-bit Montgomery, integer only", "l2": "inputs (%.1f GB) exceed L2" % (96.0 * n / 1e9),
                   "parallelism": "point-range shards + all-gather of per-rank results" if world > 1 else "single GPU"},
        "e2e": {"value": round(e2e_value, 3), "unit": UNIT, "h2d_bytes_per_step": 96 * n, "d2h_bytes_per_step": 64,
                "ms_per_step": round(e2e_ms, 4), "host_numa_binding": bool(numa_cores)},
        "gpu_launches": int(launches),
        "clocks": clocks.window(t_w0, t_w1),
        "roofline": {"bound": "hbm", "kernel": "k_msm_accumulate<Fq>", "achieved": round(achieved, 2) if achieved else None,
                     "peak": hbm_peak, "unit": "GB/s", "frac": round(achieved / hbm_peak, 5) if achieved else None,
                     "traffic": traffic, "peak_source": peak_src, "algorithmic_bytes_per_launch": 96 * n,
                     "kernel_ms": round(acc_avg_ms, 4), "launches_per_step": int(acc_k // max(1, args.steps)),
                     "limiting_pipe": {"pipe": "fma-heavy (IMAD.WIDE)", "busy_frac": pipe_busy,
                                       "source": "ncu capture committed under profiles/ (not measured live)"},
                     "ncu_capture": capture,
                     "note": "bound by the integer multiplier pipe, not by HBM (see DESIGN.md section 4); traffic = DRAM bytes "
                             "of the accumulation of one MSM from the ncu capture; kernel time / step time = %.2f (the "
                             "sort of the next window group runs concurrently)" % (acc_avg_ms / dev_ms if dev_ms else 0)},
        # the roofline that actually binds: Montgomery multiplications per second against the measured multiplier ceiling
        # (tools/imad_bench.cu, profiles/r1_imad_microbench.txt); one XYZZ mixed add = 10 field multiplications, and the
        # accumulation performs one per (term, window): 14 windows of 18-19 bits at 2^26 (csrc/msm_impl.cuh msm_geometry)
        "compute_roofline": (lambda nw: {"unit": "G field-mul/s", "achieved": round(10.0 * nw * n / (acc_avg_ms * 1e-3) / 1e9, 2),
                                         "peak": mul_peak, "frac": round(10.0 * nw * n / (acc_avg_ms * 1e-3) / 1e9 / mul_peak, 4)
                                         if mul_peak else None, "windows": nw, "peak_source": "measured microbenchmark"})(
            msm_windows(args.log_n)) if acc_avg_ms else None,
        "kernels_ms_per_step": {k: round(v[0] / args.steps, 4) for k, v in prof.items()},
        "result_x": result["r"][:32].hex(),
    }
    if pageable:
        line["e2e"]["pageable_caller_buffers_msm_2^%d" % (pageable["terms"].bit_length() - 1)] = pageable
    clocks.stop()

    if scale_rows:
        line["scaling_rows"] = scale_rows
    if world == 1:
        if not args.no_extras:
            pts = make_points(torch, np, ctx, 0, 1 << 22, 1, device)
            sc = make_scalars(torch, 1 << 22, 0x8d313d76, device)
            try:
                line["extras"] = extras(args, torch, np, ctx, stream, device, lib, pts, sc)
            except Exception as exc:       # noqa: BLE001 -- the headline line must survive a failing side measurement
                line["extras"] = {"error": "%s: %s" % (type(exc).__name__, exc)}
        del pts, sc
        torch.cuda.empty_cache()
        pot10 = line.get("extras", {}).pop("_pot10", None)
        if not args.no_cpu:
            sys.path.insert(0, os.path.join(ROOT, "oracle"))
            if pot10:   # config 1: 2^10 challenge on ONE CPU thread (the reference's own runnable case) vs the GPU response
                import oracle as oc
                chb, key, body = pot10
                t0 = time.perf_counter()
                exp = oc.pot_transform(chb, 10, 256, be(key.tau), be(key.alpha), be(key.beta), threads=1)
                line["extras"]["pot_transform_2^10"].update({"cpu_1thread_s": round(time.perf_counter() - t0, 3),
                                                             "response_matches_oracle": exp[64:] == body})
            cores = host_cores()
            line["cpu_baseline"] = cpu_line
            line["msm_2^%d_matches_oracle" % args.log_n] = bool(cpu_line and cpu_line["result_matches_gpu"])
            if "extras" in line:    # the CPU restatement of the other paths, bounded samples, same host cores
                import numpy as np
                import oracle as oc
                other = {}
                for grp, gen, lg in ((0, G1_GEN, 15), (1, G2_GEN, 13)):
                    m = 1 << lg
                    t0 = time.perf_counter()
                    oc.batch_mul_powers(grp, gen * m, be(TAU), None, 1, 0, 1, threads=cores)
                    dtc = time.perf_counter() - t0
                    other["g%d_batch_exp_2^%d" % (grp + 1, lg)] = {"s": round(dtc, 3), "Mmul_per_s": round(m / dtc / 1e6, 4)}
                xs = np.random.default_rng(3).integers(0, 256, size=(1 << 20, 32), dtype=np.uint8)
                xs[:, 0] &= 0x1f
                xs = xs.tobytes()
                t0 = time.perf_counter()
                oc.fr_fft(xs, threads=cores)
                dtc = time.perf_counter() - t0
                other["fr_fft_2^20"] = {"s": round(dtc, 3), "Melem_per_s": round((1 << 20) / dtc / 1e6, 3)}
                line["cpu_baseline"]["other_paths"] = other
    sys.stdout.flush()
    os.dup2(real_stdout, 1)
    print(json.dumps(line), flush=True)
    if scale_failed:
        os._exit(0)                       # the other ranks have left (or are stuck in a collective): no barrier to wait in
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


def extras(args, torch, np, ctx, stream, device, lib, pts, sc):
    """The other single-GPU configs of BASELINE.json, each timed a few times after the headline run."""
    out = {}
    nb = be(0x2b5d1c3e7f9a0b4c6d8e0f1a2b3c4d5e6f708192a3b4c5d6e7f8091a2b3c4d5e % R_MOD)
    k = np.frombuffer(nb, dtype=np.uint8)
    # -- config 2: G1 MSM at 2^20 (device resident)
    m = 1 << 20
    t, _ = timed(torch, stream, lambda: ctx.msm_dev(0, pts.data_ptr(), sc.data_ptr(), m), 5)
    out["g1_msm_2^20"] = {"ms": round(t, 3), "Mscalar_mul_per_s": round(m / t / 1e3, 2)}
    # -- batch_exp (the transform / contribute hot loop), device resident, 2^22 G1 and 2^20 G2
    m = min(1 << 22, pts.numel() // 64)
    outb = torch.empty(m * 64, dtype=torch.uint8, device=device)
    ctx.batch_mul_dev(0, pts.data_ptr(), outb.data_ptr(), m, k); ctx.sync()
    ctx.profile(True)
    t, _ = timed(torch, stream, lambda: (ctx.batch_mul_dev(0, pts.data_ptr(), outb.data_ptr(), m, k), ctx.sync()), 3)
    kms, kk = ctx.profile_read(lib.PROF_BATCH_MUL)
    out["g1_batch_exp_2^%d" % (m.bit_length() - 1)] = {"ms": round(t, 3), "Mmul_per_s": round(m / t / 1e3, 2),
                                                      "k_batch_mul_ms": round(kms / max(1, kk), 3),
                                                      "hbm_GBps_algorithmic": round(128.0 * m / (t * 1e-3) / 1e9, 2)}
    ctx.profile(False)
    m2 = 1 << 20
    p2 = make_points(torch, np, ctx, 1, m2, 1, device)
    o2 = torch.empty(m2 * 128, dtype=torch.uint8, device=device)
    ctx.batch_mul_dev(1, p2.data_ptr(), o2.data_ptr(), 1024, k); ctx.sync()
    t, _ = timed(torch, stream, lambda: (ctx.batch_mul_dev(1, p2.data_ptr(), o2.data_ptr(), m2, k), ctx.sync()), 2)
    out["g2_batch_exp_2^20"] = {"ms": round(t, 3), "Mmul_per_s": round(m2 / t / 1e3, 2), "probe_verdict": ctx.g2_probe_stats()[1],
                                "note": "default: batch subgroup probe on the device (8 random window sums, [r]W = O), then the "
                                        "endomorphism split; verdict 0 = proven"}
    ctx.batch_mul_dev(1, p2.data_ptr(), o2.data_ptr(), 1024, k, flags=lib.G2_EXACT); ctx.sync()
    t, _ = timed(torch, stream, lambda: (ctx.batch_mul_dev(1, p2.data_ptr(), o2.data_ptr(), m2, k, flags=lib.G2_EXACT), ctx.sync()), 2)
    out["g2_batch_exp_2^20_exact_flag"] = {"ms": round(t, 3), "Mmul_per_s": round(m2 / t / 1e3, 2),
                                           "note": "P2B_G2_EXACT: no probe, no split (the path a batch with a point outside the subgroup takes)"}
    ctx.batch_mul_dev(1, p2.data_ptr(), o2.data_ptr(), 1024, k, flags=lib.G2_SUBGROUP); ctx.sync()
    t, _ = timed(torch, stream, lambda: (ctx.batch_mul_dev(1, p2.data_ptr(), o2.data_ptr(), m2, k, flags=lib.G2_SUBGROUP), ctx.sync()), 2)
    out["g2_batch_exp_2^20_subgroup_flag"] = {"ms": round(t, 3), "Mmul_per_s": round(m2 / t / 1e3, 2),
                                              "note": "opt-in endomorphism split (P2B_G2_SUBGROUP)"}
    # -- config 5 ingredient: G2 MSM at 2^20
    ctx.msm_dev(1, p2.data_ptr(), sc.data_ptr(), m2)
    t, _ = timed(torch, stream, lambda: ctx.msm_dev(1, p2.data_ptr(), sc.data_ptr(), m2), 3)
    out["g2_msm_2^20"] = {"ms": round(t, 3), "Mscalar_mul_per_s": round(m2 / t / 1e3, 2)}
    del p2, o2
    # -- config 3: phase-2 contribute on a synthetic 2^20-constraint MPCParameters (h = 2^20 - 1, l = 2^20 points), host in / host out
    nh, nl = (1 << 20) - 1, 1 << 20
    if pts.numel() >= (nh + nl) * 64:
        hl = pts[: (nh + nl) * 64].cpu().numpy().tobytes()
        g1 = lambda i: hl[64 * i: 64 * i + 64]
        g2 = G2_GEN
        import hashlib
        import struct
        params = bytearray()
        params += g1(0) + g1(1) + g2 + g2 + G1_GEN + g2                      # alpha_g1 beta_g1 beta_g2 gamma_g2 delta_g1 delta_g2
        params += struct.pack(">I", 2) + g1(2) + g1(3)                        # ic
        params += struct.pack(">I", nh) + hl[: nh * 64]                       # h
        params += struct.pack(">I", nl) + hl[nh * 64: (nh + nl) * 64]         # l
        params += struct.pack(">I", 16) + hl[: 16 * 64]                       # a
        params += struct.pack(">I", 16) + hl[: 16 * 64]                       # b_g1
        params += struct.pack(">I", 16) + g2 * 16                             # b_g2
        params += hashlib.blake2b(b"bench").digest() + struct.pack(">I", 0)
        pin = torch.empty(len(params), dtype=torch.uint8, pin_memory=True)
        pin.copy_(torch.frombuffer(params, dtype=torch.uint8))
        pout = torch.empty(len(params) + 384, dtype=torch.uint8, pin_memory=True)
        times = []
        s_g1 = np.frombuffer(g1(5), dtype=np.uint8)
        r_g2 = np.frombuffer(lib.hash_to_g2(ctx.phase2_transcript(pin.numpy(), k, s_g1)), dtype=np.uint8)   # keypair(): r = hash_to_g2(transcript)
        for _ in range(3):
            torch.cuda.synchronize()
            t0 = time.perf_counter()
            _, h = ctx.phase2_contribute(pin.numpy(), k, s_g1, r_g2, out=pout.numpy())
            times.append(time.perf_counter() - t0)
        out["phase2_contribute_2^20"] = {"wall_s": round(min(times), 4), "points": nh + nl, "params_bytes": len(params),
                                         "contribution_hash": h.hex()[:32]}
        # -- next row: verify_contribution of that contribution: merge_pairs over H and L = four 2^20-term
        #    MSMs on the GPU, same_ratio pairings on the host
        from phase2_bn254_b200.phase2 import MPCParameters, verify_contribution
        t0 = time.perf_counter()
        vh = verify_contribution(MPCParameters(pin.numpy()), MPCParameters(pout.numpy()), ctx=ctx, rng=np.random.default_rng(7))
        out["phase2_verify_contribution_2^20"] = {"wall_s": round(time.perf_counter() - t0, 4), "accepted": bool(vh == h)}
        # -- next row: the QAP evaluation of MPCParameters::new as a sparse group linear map: 2^20 variables
        #    (rows) with 3 random (coefficient, constraint) entries each over 2^20 Lagrange-basis points
        nv = 1 << 20
        rs = np.random.default_rng(11)
        offs = np.arange(0, 3 * nv + 1, 3, dtype=np.uint64)
        cols = rs.integers(0, nv, size=3 * nv, dtype=np.uint32)
        cf = np.frombuffer(bytearray(rs.bytes(96 * nv)), dtype=np.uint8).reshape(-1, 32)
        cf[:, 0] &= 0x1f
        bases = np.frombuffer(hl[: nv * 64], dtype=np.uint8)
        ctx.sparse_mul(0, bases, offs[:1025], cols[:3072], cf[:3072].reshape(-1))
        dt = 1e9
        for _ in range(2):                                   # the first full-size call also grows the library's scratch buffers
            t0 = time.perf_counter()
            sp = ctx.sparse_mul(0, bases, offs, cols, cf.reshape(-1))
            dt = min(dt, time.perf_counter() - t0)
        import hashlib as _hl
        out["mpc_new_sparse_mul_2^20"] = {"wall_s": round(dt, 4), "rows": nv, "entries": 3 * nv, "Mentries_per_s": round(3 * nv / dt / 1e6, 2),
                                          "out_blake2b": _hl.blake2b(sp.tobytes()).hexdigest()[:32]}
    # -- configs 1 / phase-1 hot path: BatchedAccumulator::transform on the deterministic initial challenge (all generators,
    #    new_constrained), host maps in / out, compressed response; Blake2b-512 of the response body = "response hash"
    import hashlib
    from phase2_bn254_b200.powersoftau import BatchedAccumulator, CeremonyParams, PrivateKey
    key = PrivateKey(TAU, 0x2222 * 2**190 % R_MOD, 0x3333 * 2**180 % R_MOD)
    for size in (10, 20):
        prm = CeremonyParams(size, 256)
        pl, pg1 = prm.powers_length, prm.powers_g1_length
        ch = torch.empty(prm.accumulator_size, dtype=torch.uint8, pin_memory=True)
        chn = ch.numpy()
        chn[:64] = np.frombuffer(hashlib.blake2b(b"").digest(), dtype=np.uint8)
        g1a, g2a = np.frombuffer(G1_GEN, dtype=np.uint8), np.frombuffer(G2_GEN, dtype=np.uint8)
        o = 64
        for cnt, g in ((pg1, g1a), (pl, g2a), (pl, g1a), (pl, g1a), (1, g2a)):
            chn[o:o + cnt * g.size].reshape(cnt, g.size)[:] = g
            o += cnt * g.size
        rs = torch.zeros(prm.contribution_size, dtype=torch.uint8, pin_memory=True)
        rsn = rs.numpy()
        times = []
        for _ in range(3):
            torch.cuda.synchronize()
            t0 = time.perf_counter()
            BatchedAccumulator.transform(chn, rsn, False, True, False, key, prm, ctx=ctx)
            times.append(time.perf_counter() - t0)
        end = prm.contribution_size - prm.public_key_size
        out["pot_transform_2^%d" % size] = {"wall_s": round(min(times), 4), "g1_points": pg1 + 2 * pl, "g2_points": pl + 1,
                                            "challenge_bytes": prm.accumulator_size,
                                            "response_body_blake2b": hashlib.blake2b(rsn[64:end].tobytes()).hexdigest()[:32]}
        if size == 10:
            out["_pot10"] = (chn.tobytes(), key, rsn[64:end].tobytes())
        if size == 20:
            # the same call the way compute_constrained makes it: memory maps of the challenge and response FILES (pageable)
            import shutil
            import tempfile
            d = tempfile.mkdtemp(prefix="p2b_bench_", dir=args.mmap_dir or ("/dev/shm" if os.path.isdir("/dev/shm") else None))
            try:
                chn.tofile(os.path.join(d, "challenge"))
                with open(os.path.join(d, "response"), "wb") as f:
                    f.truncate(prm.contribution_size)
                cm = np.memmap(os.path.join(d, "challenge"), dtype=np.uint8, mode="r")
                rm = np.memmap(os.path.join(d, "response"), dtype=np.uint8, mode="r+")
                tm = []
                for _ in range(3):
                    torch.cuda.synchronize()
                    t0 = time.perf_counter()
                    BatchedAccumulator.transform(cm, rm, False, True, False, key, prm, ctx=ctx)
                    tm.append(time.perf_counter() - t0)
                out["pot_transform_2^20"].update({"mmap_files_wall_s": round(min(tm), 4),
                                                  "mmap_same_bytes": bool(np.array_equal(rm[64:end], rsn[64:end]))})
                del cm, rm
            finally:
                shutil.rmtree(d, ignore_errors=True)
            # -- the whole participant step as the compute_constrained binary performs it (hash challenge, key pair, transform,
            #    public key, hash response): the two BLAKE2b file hashes are sequential host work at ~1 GB/s; the challenge hash
            #    runs on a host thread next to the GPU transform (only the public key needs it), the reference's order beside it
            from phase2_bn254_b200.powersoftau import contribute_challenge
            step = {}
            try:
                for name, ov in (("overlapped", True), ("reference_order", False)):
                    rs2 = np.zeros(prm.contribution_size, dtype=np.uint8)
                    torch.cuda.synchronize()
                    t0 = time.perf_counter()
                    _, rh, _ = contribute_challenge(chn, rs2, lib.ChaChaRng([9] * 8), prm, ctx=ctx, overlap=ov)
                    step[name + "_wall_s"] = round(time.perf_counter() - t0, 4)
                    step.setdefault("response_hashes", []).append(rh.hex()[:32])
                t0 = time.perf_counter()
                hashlib.blake2b(chn).digest()
                step["challenge_blake2b_alone_s"] = round(time.perf_counter() - t0, 4)
                step["same_response"] = len(set(step.pop("response_hashes"))) == 1
            except Exception as exc:       # noqa: BLE001
                step = {"error": "%s: %s" % (type(exc).__name__, exc)}
            out["pot_contribute_step_2^20"] = step
            # -- next row: verify_transformation of that response (compressed) against the challenge: per chunk
            #    of 2^18 powers eight Pippenger MSMs on the GPU (power_pairs over tau_g1, tau_g2, alpha_g1, beta_g1), the
            #    same_ratio pairings on host threads
            from phase2_bn254_b200.powersoftau import public_key_for, verify_transformation
            digest = hashlib.blake2b(b"bench digest").digest()
            pub = public_key_for(key, lib.ChaChaRng([7] * 8), digest)
            t0 = time.perf_counter()
            ok = verify_transformation(chn, rsn, pub, digest, False, True, False, False, CeremonyParams(size, 1 << 18), ctx=ctx,
                                       rng=np.random.default_rng(3))
            out["pot_verify_transformation_2^20"] = {"wall_s": round(time.perf_counter() - t0, 4), "accepted": bool(ok),
                                                     "chunk": 1 << 18}
            # -- next row: prepare_phase2 for m = 16 from that compressed response: 3 G1 + 1 G2 group iFFTs
            #    of 2^16 points (d/2 log d scalar multiplications each) + the H query
            from phase2_bn254_b200.powersoftau import prepare_phase2
            times = []
            for _ in range(2):
                torch.cuda.synchronize()
                t0 = time.perf_counter()
                img = prepare_phase2(ctx, rsn[:end], prm, 16, input_is_compressed=True)
                times.append(time.perf_counter() - t0)
            out["prepare_phase2_m16"] = {"wall_s": round(min(times), 4), "file_bytes": int(img.size),
                                         "scalar_muls": 4 * (1 << 15) * 15 + 4 * (1 << 16),
                                         "file_blake2b": hashlib.blake2b(img.tobytes()).hexdigest()[:32]}
        del ch, rs
    # -- config 4: Fr FFT / iFFT at 2^24, round trip bit-exact
    lf = 24
    x = make_scalars(torch, 1 << lf, 0x3237db17, device)
    y = x.clone()
    ctx.fr_fft_dev(y.data_ptr(), lf, False, False); ctx.sync()
    ctx.profile(True)
    tf, _ = timed(torch, stream, lambda: (ctx.fr_fft_dev(y.data_ptr(), lf, True, False), ctx.fr_fft_dev(y.data_ptr(), lf, False, False)), 2)
    pms, pk = ctx.profile_read(lib.PROF_FFT_PASS)
    ctx.profile(False)
    ctx.fr_fft_dev(y.data_ptr(), lf, True, False); ctx.sync()
    ok = bool(torch.equal(x, y))
    out["fr_fft_2^24"] = {"ms_per_transform": round(tf / 2, 3), "Melem_per_s": round((1 << lf) / (tf / 2) / 1e3, 1),
                          "hbm_GBps_algorithmic": round(64.0 * (1 << lf) / (tf / 2 * 1e-3) / 1e9, 1),
                          "pass_ms": round(pms / max(1, pk), 3), "round_trip_bit_exact": ok}
    return out


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--log-n", type=int, default=26, help="log2 of the MSM terms per GPU")
    ap.add_argument("--ref-log-n", type=int, default=0, help="--impl reference: log2 of the terms (default: --log-n, the same config)")
    ap.add_argument("--ref-budget-s", type=float, default=120.0, help="--impl reference: stop starting new full-size steps after this long")
    ap.add_argument("--no-extras", action="store_true")
    ap.add_argument("--mmap-dir", default=None, help="directory for the memory-mapped input files of the pageable-buffer extras")
    ap.add_argument("--no-cpu", action="store_true")
    ap.add_argument("--config5-log-n", type=int, default=25, help="scaling rows: log2 of the G1 / G2 MSM terms per GPU (config 5)")
    ap.add_argument("--sharded-transform-size", type=int, default=22, help="scaling rows: log2 of the powers of the sharded transform")
    ap.add_argument("--scaling-rows-timeout", type=int, default=420, help="seconds after which the scaling rows are abandoned")
    ap.add_argument("--prepare-m", type=int, default=20, help="scaling rows: m of the sharded prepare_phase2")
    ap.add_argument("--max-contribute-log", type=int, default=26, help="scaling rows: largest log2(constraints) of the sharded contribute")
    args = ap.parse_args()
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if args.impl == "reference":
        sys.path.insert(0, os.path.join(ROOT, "oracle"))
        run_reference(args, rank)
        return
    if world != args.gpus and rank == 0:
        print("note: --gpus %d but WORLD_SIZE=%d; running %d rank(s)" % (args.gpus, world, world), file=sys.stderr)
    run_ours(args, rank, world, local_rank)


if __name__ == "__main__":
    main()
