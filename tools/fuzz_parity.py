"""One-off randomized differential run: GPU (through the C ABI) vs the CPU oracle on larger random samples than the
test suite uses.  Prints one line per family; exits non-zero on the first mismatch."""
import os, random, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "oracle")); sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np
import oracle as oc
from phase2_bn254_b200 import lib
from util import G1_GEN, G2_GEN, R_MOD, be

ctx = lib.Context(0)
seed = int(sys.argv[1]) if len(sys.argv) > 1 else 20261017
rng = random.Random(seed)
T = os.cpu_count() or 8
t0 = time.time()

def scalars(n):
    out = []
    for _ in range(n):
        c = rng.random()
        if c < 0.02: v = rng.choice([0, 1, 2, R_MOD - 1, R_MOD - 2, 1 << 253, (1 << 128) - 1, 1 << 128, 1 << 127])
        elif c < 0.06: v = rng.randrange(1 << rng.randrange(1, 254))
        elif c < 0.08: v = R_MOD - rng.randrange(1, 1 << 64)
        else: v = rng.randrange(R_MOD)
        out.append(be(v % R_MOD))
    return b"".join(out)

for group, n in ((0, 200000), (1, 40000)):
    gen = G2_GEN if group else G1_GEN
    pts = oc.batch_mul(group, gen * n, scalars(n), threads=T)          # random points (some at infinity from k = 0)
    sc = scalars(n)
    for ie, oe in ((0, 0), (0, 1)):
        exp = oc.batch_mul(group, pts, sc, ie, oe, threads=T)
        got = ctx.batch_mul(group, pts, sc, ie, oe).tobytes()
        assert got == exp, ("batch_mul", group, ie, oe)
    comp = oc.batch_mul(group, pts, be(1), 0, 1, threads=T)
    assert ctx.batch_mul(group, comp, sc, 1, 1).tobytes() == oc.batch_mul(group, pts, sc, 0, 1, threads=T), ("compressed in", group)
    k = scalars(1)
    assert ctx.batch_mul(group, pts, k).tobytes() == oc.batch_mul(group, pts, k, threads=T), ("broadcast", group)
    tau, coeff = be(rng.randrange(2, R_MOD)), be(rng.randrange(2, R_MOD))
    start = rng.randrange(1 << 27)
    m = 20000
    nz = oc.batch_mul(group, gen * m, b"".join(be(rng.randrange(1, R_MOD)) for _ in range(m)), threads=T)   # no infinity: phase-1 rejects it
    assert ctx.batch_mul_powers(group, nz, tau, coeff, start, 0, 1).tobytes() == \
        oc.batch_mul_powers(group, nz, tau, coeff, start, 0, 1, threads=T), ("powers", group)
    print("G%d batch_exp: %d points x {per-point, broadcast, compressed in/out, tau-powers}: identical (%.0f s)" % (group + 1, n, time.time() - t0), flush=True)
    # MSM at random sizes on slices of the same points
    size = 128 if group else 64
    for _ in range(12):
        m = rng.randrange(1, 20000 if group == 0 else 6000)
        o = rng.randrange(0, n - m)
        p, s = pts[o * size:(o + m) * size], sc[o * 32:(o + m) * 32]
        assert ctx.msm(group, p, s) == oc.msm(group, p, s, threads=T), ("msm", group, m, o)
    print("G%d msm: 12 random sizes < %d: identical (%.0f s)" % (group + 1, 20000 if group == 0 else 6000, time.time() - t0), flush=True)
for log_n in range(0, 19):
    x = b"".join(be(rng.randrange(R_MOD)) for _ in range(1 << log_n)) if log_n <= 12 else \
        np.random.default_rng(seed + log_n).integers(0, 256, size=(1 << log_n, 32), dtype=np.uint8)
    if not isinstance(x, bytes):
        x[:, 0] &= 0x1f
        x = x.tobytes()
    for inv, cos in ((0, 0), (1, 0), (0, 1), (1, 1)):
        assert ctx.fr_fft(x, inv, cos).tobytes() == oc.fr_fft(x, inv, cos, threads=T), ("fft", log_n, inv, cos)
print("Fr fft/ifft/coset variants, log n = 0..18: identical (%.0f s)" % (time.time() - t0), flush=True)
print("fuzz ok, seed", seed)
