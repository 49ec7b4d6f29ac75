"""MSM tuning probe (not the bench): device-resident G1 MSM at a few sizes for every accumulate variant / window width.
   python tools/msm_variants.py            -> spawns one child per configuration (the variant is read once per process)
   python tools/msm_variants.py child LG.. -> one configuration (env P2B_ACC_VARIANT, P2B_MSM_C)"""
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def child(sizes):
    import numpy as np
    import torch
    from phase2_bn254_b200 import lib
    R = 21888242871839275222246405745257275088548364400416034343698204186575808495617
    be = lambda v: np.frombuffer(int(v).to_bytes(32, "big"), dtype=np.uint8)
    G1 = (1).to_bytes(32, "big") + (2).to_bytes(32, "big")
    ctx = lib.Context(0)
    stream = torch.cuda.ExternalStream(ctx.stream)
    nmax = 1 << max(sizes)
    src = torch.from_numpy(np.frombuffer(G1, dtype=np.uint8).copy()).cuda().repeat(nmax)
    pts = torch.empty(nmax * 64, dtype=torch.uint8, device="cuda")
    torch.cuda.synchronize()
    ctx.batch_mul_powers_dev(0, src.data_ptr(), pts.data_ptr(), nmax, be(0x1234567 ** 7 % R), None, 1)
    ctx.sync()
    del src
    g = torch.Generator(device="cuda")
    g.manual_seed(1)
    sc = torch.randint(0, 256, (nmax, 32), dtype=torch.uint8, device="cuda", generator=g)
    sc[:, 0] &= 0x1f
    torch.cuda.synchronize()
    tag = "variant=%s c=%s" % (os.environ.get("P2B_ACC_VARIANT", "default"), os.environ.get("P2B_MSM_C", "auto"))
    for lg in sizes:
        n = 1 << lg
        res = None
        for _ in range(2):
            res = ctx.msm_dev(0, pts.data_ptr(), sc.data_ptr(), n)
        ctx.profile(True)
        reps = 3 if lg >= 24 else 10
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        torch.cuda.synchronize()
        e0.record(stream)
        for _ in range(reps):
            res = ctx.msm_dev(0, pts.data_ptr(), sc.data_ptr(), n)
        e1.record(stream)
        torch.cuda.synchronize()
        t = e0.elapsed_time(e1) / reps
        acc = ctx.profile_read(lib.PROF_MSM_ACCUMULATE)[0] / reps
        srt = ctx.profile_read(lib.PROF_MSM_SORT)[0] / reps
        red = ctx.profile_read(lib.PROF_MSM_REDUCE)[0] / reps
        ctx.profile(False)
        print("%s 2^%d: %8.3f ms  %7.1f M/s  acc %.3f sort %.3f reduce %.3f  x=%s" % (tag, lg, t, n / t / 1e3, acc, srt, red, res[:8].hex()),
              flush=True)


if __name__ == "__main__":
    if len(sys.argv) > 1 and sys.argv[1] == "child":
        child([int(a) for a in sys.argv[2:]])
    else:
        sizes = sys.argv[1:] or ["20", "22", "26"]
        configs = [("0", None), ("1", None), ("2", None), ("3", None)]
        for v in os.environ.get("EXTRA_C", "").split(","):
            if v:
                configs += [("1", v), ("2", v)]
        for variant, c in configs:
            env = dict(os.environ, P2B_ACC_VARIANT=variant)
            if c:
                env["P2B_MSM_C"] = c
            subprocess.run([sys.executable, os.path.abspath(__file__), "child"] + sizes, env=env)
