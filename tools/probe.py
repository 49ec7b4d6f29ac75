"""Quick device-side throughput probe (not the bench): batch-mul G1/G2 at a few sizes."""
import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from phase2_bn254_b200 import lib

R = 21888242871839275222246405745257275088548364400416034343698204186575808495617
be = lambda v: np.frombuffer(int(v).to_bytes(32, "big"), dtype=np.uint8)
G1 = (1).to_bytes(32, "big") + (2).to_bytes(32, "big")
G2 = b"".join(v.to_bytes(32, "big") for v in (
    11559732032986387107991004021392285783925812861821192530917403151452391805634,
    10857046999023057135944570762232829481370756359578518086990519993285655852781,
    4082367875863433681332203403145435568316851327593401208105741076214120093531,
    8495653923123431417604973247489272438418190587263600148770280649306958101930))

ctx = lib.Context(0)
stream = torch.cuda.ExternalStream(ctx.stream)
sizes = [int(s) for s in sys.argv[1:]] or [16, 20]
for group, gen in ((0, G1), (1, G2)):
    for lg in sizes:
        n = 1 << lg
        esz = len(gen)
        src = torch.from_numpy(np.frombuffer(gen, dtype=np.uint8).copy()).cuda().repeat(n)
        pts = torch.empty(n * esz, dtype=torch.uint8, device="cuda")
        out = torch.empty(n * esz, dtype=torch.uint8, device="cuda")
        torch.cuda.synchronize()
        ctx.batch_mul_powers_dev(group, src.data_ptr(), pts.data_ptr(), n, be(0x1234567 ** 7 % R), None, 1)
        ctx.sync()
        for mode in ("powers", "broadcast"):
            ts = []
            for it in range(4):
                e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                e0.record(stream)
                if mode == "powers":
                    ctx.batch_mul_powers_dev(group, pts.data_ptr(), out.data_ptr(), n, be(0x7654321 ** 5 % R), be(R - 7), 5, 0, 1)
                else:
                    ctx.batch_mul_dev(group, pts.data_ptr(), out.data_ptr(), n, be(0xabcdef ** 9 % R))
                e1.record(stream)
                ctx.sync()
                ts.append(e0.elapsed_time(e1))
            t = min(ts[1:])
            print("G%d 2^%d %-9s %8.3f ms  %8.3f Mmul/s" % (group + 1, lg, mode, t, n / t / 1e3), flush=True)

# ---- MSM probe
def msm_probe(lg, group=0):
    n = 1 << lg
    gen = G2 if group else G1
    esz = len(gen)
    src = torch.from_numpy(np.frombuffer(gen, dtype=np.uint8).copy()).cuda().repeat(n)
    pts = torch.empty(n * esz, dtype=torch.uint8, device="cuda")
    torch.cuda.synchronize()
    ctx.batch_mul_powers_dev(group, src.data_ptr(), pts.data_ptr(), n, be(0x1234567 ** 7 % R), None, 1)
    ctx.sync()
    g = torch.Generator(device="cuda"); g.manual_seed(lg)
    sc = torch.randint(0, 256, (n, 32), dtype=torch.uint8, device="cuda", generator=g)
    sc[:, 0] &= 0x1f
    if os.environ.get("SKEW"):          # every scalar equal: one bucket per window receives all n terms
        sc[:] = sc[0].clone()
    torch.cuda.synchronize()
    ts = []
    for it in range(4):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(stream)
        r = ctx.msm_dev(group, pts.data_ptr(), sc.data_ptr(), n)
        e1.record(stream)
        torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1))
    t = min(ts[1:])
    print("MSM G%d 2^%d %8.3f ms  %8.3f Mterm/s  result %s" % (group + 1, lg, t, n / t / 1e3, r[:8].hex()), flush=True)

if os.environ.get("MSM"):
    for lg in [int(x) for x in os.environ["MSM"].split(",")]:
        msm_probe(lg, 0)
    if os.environ.get("MSM_G2"):
        for lg in [int(x) for x in os.environ["MSM_G2"].split(",")]:
            msm_probe(lg, 1)
