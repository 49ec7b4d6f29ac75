"""One-off: wall time of BatchedAccumulator::transform on the initial challenge at larger sizes (host maps in / out)."""
import hashlib, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import bench
from phase2_bn254_b200 import lib
from phase2_bn254_b200.powersoftau import BatchedAccumulator, CeremonyParams, PrivateKey

ctx = lib.Context(0)
key = PrivateKey(bench.TAU, 0x2222 * 2**190 % bench.R_MOD, 0x3333 * 2**180 % bench.R_MOD)
for size in [int(a) for a in sys.argv[1:]] or [22]:
    prm = CeremonyParams(size, 256)
    ch = torch.empty(prm.accumulator_size, dtype=torch.uint8, pin_memory=True).numpy()
    ch[:64] = np.frombuffer(hashlib.blake2b(b"").digest(), dtype=np.uint8)
    o = 64
    g1a, g2a = np.frombuffer(bench.G1_GEN, dtype=np.uint8), np.frombuffer(bench.G2_GEN, dtype=np.uint8)
    for cnt, g in ((prm.powers_g1_length, g1a), (prm.powers_length, g2a), (prm.powers_length, g1a), (prm.powers_length, g1a), (1, g2a)):
        ch[o:o + cnt * g.size].reshape(cnt, g.size)[:] = g
        o += cnt * g.size
    rs = torch.zeros(prm.contribution_size, dtype=torch.uint8, pin_memory=True).numpy()
    end = prm.contribution_size - prm.public_key_size
    for flag in (False, True):
        ts = []
        for _ in range(2):
            torch.cuda.synchronize(); t0 = time.perf_counter()
            BatchedAccumulator.transform(ch, rs, False, True, False, key, prm, ctx=ctx, g2_in_subgroup=flag)
            ts.append(time.perf_counter() - t0)
        t0 = time.perf_counter(); h = hashlib.blake2b(rs[64:end].tobytes()).hexdigest()[:32]; th = time.perf_counter() - t0
        print("transform 2^%d (%d G1 + %d G2 points, challenge %.2f GB) g2_subgroup_flag=%s: %.3f s; response body blake2b %s (hashing %.2f s on 1 host thread)"
              % (size, prm.powers_g1_length + 2 * prm.powers_length, prm.powers_length + 1, prm.accumulator_size / 1e9, flag, min(ts), h, th), flush=True)
