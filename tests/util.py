"""Shared helpers for the parity tests: seeded synthetic inputs built with the oracle."""
import random

R_MOD = 21888242871839275222246405745257275088548364400416034343698204186575808495617
Q_MOD = 21888242871839275222246405745257275088696311157297823662689037894645226208583
LAMBDA = 0xb3c4d79d41a917585bfc41088d8daaa78b17ea66b99c90dd
G1_GEN = (1).to_bytes(32, "big") + (2).to_bytes(32, "big")
G2_GEN = b"".join(v.to_bytes(32, "big") for v in (
    11559732032986387107991004021392285783925812861821192530917403151452391805634,
    10857046999023057135944570762232829481370756359578518086990519993285655852781,
    4082367875863433681332203403145435568316851327593401208105741076214120093531,
    8495653923123431417604973247489272438418190587263600148770280649306958101930))


def be(x):
    return int(x).to_bytes(32, "big")


EDGE_SCALARS = [0, 1, 2, 3, 15, 16, 17, 31, 32, R_MOD - 1, R_MOD - 2, LAMBDA, LAMBDA - 1, LAMBDA + 1,
                R_MOD - LAMBDA, 1 << 253, (1 << 253) + 1, 1 << 128, (1 << 128) - 1, (1 << 127) + 1]


def random_points(oc, group, n, seed, threads=8):
    """n pseudo-random points [h_i]G (uncompressed wire) via the oracle's batch path."""
    rng = random.Random(seed)
    gen = G2_GEN if group else G1_GEN
    ks = b"".join(be(rng.randrange(1, R_MOD)) for _ in range(n))
    return oc.batch_mul(group, gen * n, ks, threads=threads)


def random_scalars(n, seed):
    rng = random.Random(seed)
    return b"".join(be(rng.randrange(R_MOD)) for _ in range(n))


# ---- ChaCha20 (RFC 7539 block function) in python: the reference for the coefficients the library generates on the device
def _rotl(v, k):
    return ((v << k) | (v >> (32 - k))) & 0xffffffff


def chacha20_block(key_words, w12, w13, w14, w15):
    """64 keystream bytes for the given key (8 little-endian words) and the last four state words."""
    st = [0x61707865, 0x3320646e, 0x79622d32, 0x6b206574] + list(key_words) + [w12, w13, w14, w15]
    x = list(st)

    def qr(a, b, c, d):
        x[a] = (x[a] + x[b]) & 0xffffffff; x[d] = _rotl(x[d] ^ x[a], 16)
        x[c] = (x[c] + x[d]) & 0xffffffff; x[b] = _rotl(x[b] ^ x[c], 12)
        x[a] = (x[a] + x[b]) & 0xffffffff; x[d] = _rotl(x[d] ^ x[a], 8)
        x[c] = (x[c] + x[d]) & 0xffffffff; x[b] = _rotl(x[b] ^ x[c], 7)

    for _ in range(10):
        qr(0, 4, 8, 12); qr(1, 5, 9, 13); qr(2, 6, 10, 14); qr(3, 7, 11, 15)
        qr(0, 5, 10, 15); qr(1, 6, 11, 12); qr(2, 7, 8, 13); qr(3, 4, 9, 14)
    return b"".join(((a + b) & 0xffffffff).to_bytes(4, "little") for a, b in zip(x, st))


def device_scalars(seed, n, bits=253, first=0):
    """What p2b_random_scalars / the seed-driven msm_pair use (include/p2b.h): coefficient i = keystream bytes
    [32 i, 32 i + 32) -- block counter i // 2 in words 12-13 -- as a big-endian integer, cleared above `bits` bits."""
    key = [int.from_bytes(seed[4 * i: 4 * i + 4], "little") for i in range(8)]
    out = bytearray()
    blocks = {}
    for i in range(first, first + n):
        b = i // 2
        if b not in blocks:
            blocks[b] = chacha20_block(key, b & 0xffffffff, b >> 32, 0, 0)
        v = int.from_bytes(blocks[b][32 * (i % 2): 32 * (i % 2) + 32], "big") & ((1 << bits) - 1)
        out += v.to_bytes(32, "big")
    return bytes(out)


class OracleCtx:
    """Duck-typed stand-in for lib.Context in GPU-less tests of the host-side mirrors: the Context calls they make, answered
    by the CPU oracle (and, for the device-generated verifier coefficients, by the python ChaCha20 above)."""

    def __init__(self, oc, threads=4):
        self.oc, self.threads = oc, threads

    def _wrap(self, fn, *a, **k):
        from phase2_bn254_b200 import lib
        try:
            return fn(*a, **k)
        except self.oc.OracleError as e:
            raise lib.P2BError(e.code, "oracle", e.index, e.sub)

    def msm(self, group, points, scalars):
        import numpy as np
        return self._wrap(self.oc.msm, group, bytes(np.asarray(points)), bytes(np.asarray(scalars)), threads=self.threads)

    def sum_points(self, group, points):
        return self.oc.sum_points(group, bytes(points))

    def recode(self, group, points, in_enc, out_enc, flags=0, out=None):
        import numpy as np
        from phase2_bn254_b200 import lib
        res = self._wrap(self.oc.batch_mul, group, bytes(np.asarray(points)), be(1), in_enc, out_enc, bool(flags & lib.CHECK_INPUT),
                         bool(flags & lib.REJECT_INFINITY), threads=self.threads)
        return np.frombuffer(res, dtype=np.uint8)

    def validate(self, group, points, in_enc=0, flags=1):
        self.recode(group, points, in_enc, 0, flags)

    def _coeffs(self, n, scalars, seed, bits):
        import numpy as np
        return bytes(np.asarray(scalars)) if scalars is not None else device_scalars(bytes(seed), n, bits)

    def msm_pair(self, group, points_a, points_b, scalars=None, seed=None, scalar_bits=0, in_enc=0, flags=0):
        a = self.recode(group, points_a, in_enc, 0, flags).tobytes()
        b = self.recode(group, points_b, in_enc, 0, flags).tobytes()
        k = self._coeffs(len(a) // (128 if group else 64), scalars, seed, scalar_bits)
        return self.msm(group, a, k), self.msm(group, b, k)

    def power_pairs(self, group, points, scalars=None, seed=None, scalar_bits=0, in_enc=0, flags=0):
        size = 128 if group else 64
        v = self.recode(group, points, in_enc, 0, flags).tobytes()
        n = len(v) // size
        k = self._coeffs(n - 1, scalars, seed, scalar_bits)
        return self.msm(group, v[: (n - 1) * size], k), self.msm(group, v[size:], k)

    def gfft_stage(self, group, a, b, w=None, start=0, in_enc=0, out_enc=0, flags=0, want_sum=True, want_diff=True):
        import numpy as np
        size = 128 if group else 64
        a, b = bytes(np.asarray(a)), bytes(np.asarray(b))
        n = len(a) // size
        negb = self.oc.batch_mul(group, b, be(R_MOD - 1), threads=self.threads)
        pair = lambda x, y, i: x[i * size: (i + 1) * size] + y[i * size: (i + 1) * size]
        sums = b"".join(self.oc.sum_points(group, pair(a, b, i)) for i in range(n)) if want_sum else None
        diffs = None
        if want_diff:
            diffs = b"".join(self.oc.sum_points(group, pair(a, negb, i)) for i in range(n))
            if w is not None:
                wv = int.from_bytes(bytes(np.asarray(w)), "big")
                ks = b"".join(be(pow(wv, start + i, R_MOD)) for i in range(n))
                diffs = self.oc.batch_mul(group, diffs, ks, threads=self.threads)
        f = lambda x: np.frombuffer(x, dtype=np.uint8) if x is not None else None
        return f(sums), f(diffs)

    def group_fft_scaled(self, group, points, inverse, total_log_d, in_enc=0, out_enc=0, flags=0):
        """The transform by its definition X[k] = scale * sum_n w^(n k) x[n], one oracle MSM per output."""
        import numpy as np
        from phase2_bn254_b200 import lib
        size = 128 if group else 64
        pts = bytes(np.asarray(points))
        d = len(pts) // size
        w = lib.root_of_unity(d.bit_length() - 1, inverse)
        scale = pow(pow(2, total_log_d, R_MOD), -1, R_MOD) if inverse else 1
        out = b"".join(self.oc.msm(group, pts, b"".join(be(pow(w, n * k, R_MOD) * scale % R_MOD) for n in range(d)), threads=self.threads)
                       for k in range(d))
        return np.frombuffer(out, dtype=np.uint8)

    def group_fft(self, group, points, inverse=False, in_enc=0, out_enc=0, flags=0):
        import numpy as np
        d = len(bytes(np.asarray(points))) // (128 if group else 64)
        return self.group_fft_scaled(group, points, inverse, d.bit_length() - 1)


# ---- points of the twist E'(Fq2) OUTSIDE the order-r subgroup (the reference decodes them without complaint, ec.rs:1145-1213)
G2_COFACTOR = 2 * Q_MOD - R_MOD                       # ec.rs:1347-1357; = 10069 * 5864401 * 1875725156269 * (a 177-bit prime)
G2_COFACTOR_SMALL_PRIME = 10069


def twist_points_outside_subgroup(n, seed, small_order=False):
    """n uncompressed G2 encodings of curve points that are not in the order-r subgroup, built with the big-int reference
    (oracle/bn254_ref.py).  small_order: each point is S + T with S in the subgroup and T of order 10069, the smallest
    prime factor of the cofactor -- the hardest case for a randomised membership test."""
    import bn254_ref as ref
    rng = random.Random(seed)
    out = []
    while len(out) < n:
        x = (rng.randrange(Q_MOD), rng.randrange(Q_MOD))
        y = ref.f2_sqrt(ref.f2_add(ref.f2_mul(ref.f2_sqr(x), x), ref.B_G2))
        if y is None:
            continue
        p = (x, y)
        assert ref.G2.on_curve(p)
        if small_order:
            t = ref.G2.mul(p, R_MOD * (G2_COFACTOR // G2_COFACTOR_SMALL_PRIME))
            if t is None:
                continue
            assert ref.G2.mul(t, G2_COFACTOR_SMALL_PRIME) is None
            p = ref.G2.add(ref.G2.mul(ref.G2_GEN, rng.randrange(1, R_MOD)), t)
        if ref.G2.mul(p, R_MOD) is None:
            continue
        out.append(ref.g2_encode(p, False))
    return b"".join(out)
