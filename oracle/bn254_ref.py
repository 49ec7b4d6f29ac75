"""Slow big-int restatement of the BN254 maths on the phase2/powersoftau hot path.

TEST INFRASTRUCTURE ONLY.  Nothing in the product path (phase2_bn254_b200/, the C-ABI
library) may import this module; only tests/, __graft_entry__.smoke() and bench.py's
cpu_baseline leg use oracle/ -- and only as the checker.

This file is the *second*, independent oracle: plain Python integers, affine formulas
from the curve equation, no Montgomery form, no windowing.  It exists to cross-check
oracle/p2b_oracle.c (which restates the reference's algorithms limb-for-limb) on small
cases and to derive constants.  Because every output byte on the path is the canonical
big-endian encoding of a mathematically unique group / field element,
this implementation and the reference must agree bit-for-bit.

Reference lines followed (all under /root/reference):
  moduli / generators / curve b   pairing/src/bn256/fq.rs:4-7,11-31,39-83  fr.rs:3-6
  Fq2 = Fq[u]/(u^2+1), ordering   pairing/src/bn256/fq2.rs:21-30,131-199
  point codecs and flag bits      pairing/src/bn256/ec.rs:763-946,1136-1344
  y from x / sign selection       pairing/src/bn256/ec.rs:110-131
  ceremony geometry               powersoftau/src/parameters.rs:72-120
  file positions                  powersoftau/src/batched_accumulator.rs:96-178
  transform semantics             powersoftau/src/batched_accumulator.rs:1119-1292
  initial accumulator             powersoftau/src/batched_accumulator.rs:1295-1347
  groth16 Parameters wire format  bellman/src/groth16/mod.rs:141-158,252-383
  phase2 contribute               phase2/src/parameters.rs:414-522,663-703
  phase2 PublicKey wire format    phase2/src/keypair.rs:50-105
  radix-2 domain / FFT            bellman/src/domain.rs:52-99,154-195,274-317
  Pippenger result (= plain sum)  bellman/src/multiexp.rs:330-355
"""
import hashlib
import struct

Q = 21888242871839275222246405745257275088696311157297823662689037894645226208583
R = 21888242871839275222246405745257275088548364400416034343698204186575808495617
BN_X = 4965661367192848881
assert Q == 36 * BN_X**4 + 36 * BN_X**3 + 24 * BN_X**2 + 6 * BN_X + 1
assert R == 36 * BN_X**4 + 36 * BN_X**3 + 18 * BN_X**2 + 6 * BN_X + 1

FR_S = 28                                    # fr.rs:31-34
FR_GENERATOR = 7                             # fr.rs:5
FR_ROOT_OF_UNITY = pow(FR_GENERATOR, (R - 1) >> FR_S, R)
assert FR_ROOT_OF_UNITY == 0x03ddb9f5166d18b798865ea93dd31f743215cf6dd39329c8d34f1ed960c37c9c

G1_GEN = (1, 2)                              # fq.rs:36-50 (Montgomery form of 1 and 2)
# fq.rs:54-58 (decimal), Fq2 element = (c0, c1) meaning c0 + c1*u
G2_GEN = (
    (10857046999023057135944570762232829481370756359578518086990519993285655852781,
     11559732032986387107991004021392285783925812861821192530917403151452391805634),
    (8495653923123431417604973247489272438418190587263600148770280649306958101930,
     4082367875863433681332203403145435568316851327593401208105741076214120093531),
)
B_G1 = 3                                     # fq.rs:9-16


# ----------------------------------------------------------------------------- Fq2
def f2_add(a, b): return ((a[0] + b[0]) % Q, (a[1] + b[1]) % Q)
def f2_sub(a, b): return ((a[0] - b[0]) % Q, (a[1] - b[1]) % Q)
def f2_neg(a): return ((-a[0]) % Q, (-a[1]) % Q)
def f2_mul(a, b):
    return ((a[0] * b[0] - a[1] * b[1]) % Q, (a[0] * b[1] + a[1] * b[0]) % Q)
def f2_sqr(a): return f2_mul(a, a)
def f2_inv(a):
    n = pow((a[0] * a[0] + a[1] * a[1]) % Q, -1, Q)
    return (a[0] * n % Q, (-a[1]) * n % Q)
def f2_pow(a, e):
    res = (1, 0)
    while e:
        if e & 1:
            res = f2_mul(res, a)
        a = f2_sqr(a)
        e >>= 1
    return res
def f2_gt(a, b):
    """Fq2 ordering: compare c1 first, then c0 (fq2.rs:21-30)."""
    return (a[1], a[0]) > (b[1], b[0])

B_G2 = f2_mul((3, 0), f2_inv((9, 1)))        # 3/(9+u), fq.rs:18-31


def fq_sqrt(a):
    """q = 3 mod 4 => a^((q+1)/4) (ff_derive sqrt for this modulus class)."""
    a %= Q
    y = pow(a, (Q + 1) // 4, Q)
    return y if y * y % Q == a else None


def f2_sqrt(a):
    """Any square root in Fq2 (the caller canonicalises the sign), or None."""
    if a == (0, 0):
        return (0, 0)
    # complex method: a = c0 + c1 u, norm = c0^2 + c1^2
    c0, c1 = a
    if c1 == 0:
        s = fq_sqrt(c0)
        if s is not None:
            return (s, 0)
        s = fq_sqrt((-c0) % Q)
        return (0, s) if s is not None else None
    n = fq_sqrt((c0 * c0 + c1 * c1) % Q)
    if n is None:
        return None
    inv2 = pow(2, -1, Q)
    for nn in (n, (-n) % Q):
        t = (c0 + nn) * inv2 % Q
        x0 = fq_sqrt(t)
        if x0 is not None and x0 != 0:
            x1 = c1 * pow(2 * x0, -1, Q) % Q
            if f2_sqr((x0, x1)) == a:
                return (x0, x1)
    return None


# ----------------------------------------------------------------------------- curves
class Curve:
    """Affine short-Weierstrass y^2 = x^3 + b over F (a = 0).  None = infinity."""

    def __init__(self, name, b, add, sub, mul, inv, neg, zero, three_x2_over_2y=None):
        self.name, self.b = name, b
        self.fadd, self.fsub, self.fmul, self.finv, self.fneg, self.fzero = add, sub, mul, inv, neg, zero

    def on_curve(self, p):
        if p is None:
            return True
        x, y = p
        return self.fmul(y, y) == self.fadd(self.fmul(self.fmul(x, x), x), self.b)

    def neg(self, p):
        return None if p is None else (p[0], self.fneg(p[1]))

    def add(self, p, q_):
        if p is None:
            return q_
        if q_ is None:
            return p
        x1, y1 = p
        x2, y2 = q_
        if x1 == x2:
            if y1 != y2 or y1 == self.fzero:
                return None
            x1x1 = self.fmul(x1, x1)
            num = self.fadd(self.fadd(x1x1, x1x1), x1x1)
            lam = self.fmul(num, self.finv(self.fadd(y1, y1)))
        else:
            lam = self.fmul(self.fsub(y2, y1), self.finv(self.fsub(x2, x1)))
        x3 = self.fsub(self.fsub(self.fmul(lam, lam), x1), x2)
        y3 = self.fsub(self.fmul(lam, self.fsub(x1, x3)), y1)
        return (x3, y3)

    def mul(self, p, k):
        """Plain MSB-first double-and-add on the integer k (no reduction mod r)."""
        acc = None
        for bit in bin(k)[2:] if k else "":
            acc = self.add(acc, acc)
            if bit == "1":
                acc = self.add(acc, p)
        return acc


G1 = Curve("G1", B_G1,
           lambda a, b: (a + b) % Q, lambda a, b: (a - b) % Q, lambda a, b: a * b % Q,
           lambda a: pow(a, -1, Q), lambda a: (-a) % Q, 0)
G2 = Curve("G2", B_G2, f2_add, f2_sub, f2_mul, f2_inv, f2_neg, (0, 0))
assert G1.on_curve(G1_GEN) and G2.on_curve(G2_GEN)


# ----------------------------------------------------------------------------- codecs
class DecodeError(Exception):
    pass


def _be(x):
    return x.to_bytes(32, "big")


def _fq_from(b, what):
    v = int.from_bytes(b, "big")
    if v >= Q:
        raise DecodeError("CoordinateDecodingError(%s)" % what)
    return v


def g1_encode(p, compressed):
    """ec.rs:827-843 (uncompressed), 920-945 (compressed)."""
    n = 32 if compressed else 64
    if p is None:
        return bytes([0x40]) + bytes(n - 1)
    x, y = p
    if not compressed:
        return _be(x) + _be(y)
    out = bytearray(_be(x))
    if y > (-y) % Q:
        out[0] |= 0x80
    return bytes(out)


def g1_decode(b, compressed, checked=True):
    """ec.rs:772-826, 875-919."""
    b = bytearray(b)
    if b[0] & 0x40:
        b[0] &= 0x3F
        if any(b):
            raise DecodeError("UnexpectedInformation")
        return None
    if not compressed:
        if b[0] & 0x80:
            raise DecodeError("UnexpectedInformation")
        b[0] &= 0x3F
        p = (_fq_from(b[:32], "x"), _fq_from(b[32:64], "y"))
        if checked and not G1.on_curve(p):
            raise DecodeError("NotOnCurve")
        return p
    greatest = bool(b[0] & 0x80)
    b[0] &= 0x3F
    x = _fq_from(b[:32], "x")
    y = fq_sqrt((x * x * x + B_G1) % Q)
    if y is None:
        raise DecodeError("NotOnCurve")
    negy = (-y) % Q
    return (x, y if ((y < negy) != greatest) else negy)


def g2_encode(p, compressed):
    """ec.rs:1214-1231, 1317-1343; Fq2 written c1 first then c0."""
    n = 64 if compressed else 128
    if p is None:
        return bytes([0x40]) + bytes(n - 1)
    x, y = p
    if not compressed:
        return _be(x[1]) + _be(x[0]) + _be(y[1]) + _be(y[0])
    out = bytearray(_be(x[1]) + _be(x[0]))
    if f2_gt(y, f2_neg(y)):
        out[0] |= 0x80
    return bytes(out)


def g2_decode(b, compressed, checked=True):
    """ec.rs:1145-1213, 1264-1315."""
    b = bytearray(b)
    if not compressed:
        if b[0] & 0x80:
            raise DecodeError("UnexpectedCompressionMode")
        if b[0] & 0x40:
            b[0] &= 0x3F
            if any(b):
                raise DecodeError("UnexpectedInformation")
            return None
        b[0] &= 0x3F
        xc1, xc0 = _fq_from(b[0:32], "x c1"), _fq_from(b[32:64], "x c0")
        yc1, yc0 = _fq_from(b[64:96], "y c1"), _fq_from(b[96:128], "y c0")
        p = ((xc0, xc1), (yc0, yc1))
        if checked and not G2.on_curve(p):
            raise DecodeError("NotOnCurve")
        return p
    if b[0] & 0x40:
        b[0] &= 0x3F
        if any(b):
            raise DecodeError("UnexpectedInformation")
        return None
    greatest = bool(b[0] & 0x80)
    b[0] &= 0x3F
    xc1, xc0 = _fq_from(b[0:32], "x c1"), _fq_from(b[32:64], "x c0")
    x = (xc0, xc1)
    y = f2_sqrt(f2_add(f2_mul(f2_sqr(x), x), B_G2))
    if y is None:
        raise DecodeError("NotOnCurve")
    negy = f2_neg(y)
    return (x, y if (f2_gt(negy, y) != greatest) else negy)


def fr_encode(k):
    return (k % R).to_bytes(32, "big")


# ----------------------------------------------------------------------------- phase 1
class CeremonyParams:
    """powersoftau/src/parameters.rs:72-120 for Bn256 (g1 64/32 B, g2 128/64 B)."""

    def __init__(self, size, batch_size):
        self.size, self.batch_size = size, batch_size
        self.powers_length = 1 << size
        self.powers_g1_length = (self.powers_length << 1) - 1
        self.hash_size = 64
        g1, g2, g1c, g2c = 64, 128, 32, 64
        self.accumulator_size = (self.powers_g1_length * g1 + self.powers_length * g2 +
                                 self.powers_length * g1 * 2 + g2 + self.hash_size)
        self.public_key_size = 3 * g2 + 6 * g1
        self.contribution_size = (self.powers_g1_length * g1c + self.powers_length * g2c +
                                  self.powers_length * g1c * 2 + g2c + self.hash_size +
                                  self.public_key_size)

    def position(self, index, element, compressed):
        """batched_accumulator.rs:96-178.  element in tau_g1,tau_g2,alpha_g1,beta_g1,beta_g2."""
        g1, g2 = (32, 64) if compressed else (64, 128)
        pos = {"tau_g1": g1 * index,
               "tau_g2": g1 * self.powers_g1_length + g2 * index,
               "alpha_g1": g1 * self.powers_g1_length + g2 * self.powers_length + g1 * index,
               "beta_g1": g1 * self.powers_g1_length + (g2 + g1) * self.powers_length + g1 * index,
               "beta_g2": g1 * self.powers_g1_length + (g2 + 2 * g1) * self.powers_length}[element]
        return pos + self.hash_size


def blank_hash():
    """utils.rs:138-140."""
    return hashlib.blake2b(b"").digest()


def generate_initial(params, compressed=False):
    """batched_accumulator.rs:1295-1347 + new_constrained.rs:56-63 (hash prefix)."""
    size = params.contribution_size - params.public_key_size if compressed else params.accumulator_size
    out = bytearray(size)
    out[0:64] = blank_hash()
    e1, e2 = g1_encode(G1_GEN, compressed), g2_encode(G2_GEN, compressed)
    for i in range(params.powers_g1_length):
        p = params.position(i, "tau_g1", compressed)
        out[p:p + len(e1)] = e1
    for i in range(params.powers_length):
        for el, e in (("tau_g2", e2), ("alpha_g1", e1), ("beta_g1", e1)):
            p = params.position(i, el, compressed)
            out[p:p + len(e)] = e
    p = params.position(0, "beta_g2", compressed)
    out[p:p + len(e2)] = e2
    return bytes(out)


def transform(params, challenge, tau, alpha, beta, in_compressed=False, out_compressed=True,
              checked=False):
    """Accumulator region of the response (bytes [64, len-768) of the response file, or
    [64, len) of an uncompressed new challenge), per batched_accumulator.rs:1119-1292."""
    g1i, g2i = (32, 64) if in_compressed else (64, 128)
    out = {}

    def rd1(el, i):
        p = params.position(i, el, in_compressed)
        pt = g1_decode(challenge[p:p + g1i], in_compressed, checked)
        if pt is None:
            raise DecodeError("PointAtInfinity")
        return pt

    def rd2(el, i):
        p = params.position(i, el, in_compressed)
        pt = g2_decode(challenge[p:p + g2i], in_compressed, checked)
        if pt is None:
            raise DecodeError("PointAtInfinity")
        return pt

    total = (params.contribution_size - params.public_key_size) if out_compressed \
        else params.accumulator_size
    buf = bytearray(total)
    tp = 1
    for i in range(params.powers_g1_length):
        e = g1_encode(G1.mul(rd1("tau_g1", i), tp), out_compressed)
        p = params.position(i, "tau_g1", out_compressed)
        buf[p:p + len(e)] = e
        if i < params.powers_length:
            e = g2_encode(G2.mul(rd2("tau_g2", i), tp), out_compressed)
            p = params.position(i, "tau_g2", out_compressed)
            buf[p:p + len(e)] = e
            e = g1_encode(G1.mul(rd1("alpha_g1", i), tp * alpha % R), out_compressed)
            p = params.position(i, "alpha_g1", out_compressed)
            buf[p:p + len(e)] = e
            e = g1_encode(G1.mul(rd1("beta_g1", i), tp * beta % R), out_compressed)
            p = params.position(i, "beta_g1", out_compressed)
            buf[p:p + len(e)] = e
        tp = tp * tau % R
    e = g2_encode(G2.mul(rd2("beta_g2", 0), beta), out_compressed)
    p = params.position(0, "beta_g2", out_compressed)
    buf[p:p + len(e)] = e
    return bytes(buf[64:])


# ----------------------------------------------------------------------------- phase 2
def _rd_vec(buf, off, size):
    (n,) = struct.unpack_from(">I", buf, off)
    off += 4
    return [bytes(buf[off + i * size: off + (i + 1) * size]) for i in range(n)], off + n * size


def params_parse(buf):
    """bellman/src/groth16/mod.rs:287-383 + phase2/src/parameters.rs:682-703 -> dict of raw
    encodings (no decoding)."""
    d, off = {}, 0
    for name, size in (("alpha_g1", 64), ("beta_g1", 64), ("beta_g2", 128), ("gamma_g2", 128),
                       ("delta_g1", 64), ("delta_g2", 128)):
        d[name] = bytes(buf[off:off + size])
        off += size
    d["ic"], off = _rd_vec(buf, off, 64)
    d["h"], off = _rd_vec(buf, off, 64)
    d["l"], off = _rd_vec(buf, off, 64)
    d["a"], off = _rd_vec(buf, off, 64)
    d["b_g1"], off = _rd_vec(buf, off, 64)
    d["b_g2"], off = _rd_vec(buf, off, 128)
    d["cs_hash"] = bytes(buf[off:off + 64])
    off += 64
    d["contributions"], off = _rd_vec(buf, off, 384)
    assert off == len(buf), (off, len(buf))
    return d


def params_serialize(d):
    out = bytearray()
    for name in ("alpha_g1", "beta_g1", "beta_g2", "gamma_g2", "delta_g1", "delta_g2"):
        out += d[name]
    for name in ("ic", "h", "l", "a", "b_g1", "b_g2"):
        out += struct.pack(">I", len(d[name])) + b"".join(d[name])
    out += d["cs_hash"]
    out += struct.pack(">I", len(d["contributions"])) + b"".join(d["contributions"])
    return bytes(out)


def phase2_pubkey(d, delta, s, r_g2):
    """phase2/src/parameters.rs:860-908 given the RNG-derived s (G1) and r (G2) explicitly.
    Returns the 384-byte PublicKey wire form (keypair.rs:50-62).  NOTE: in the reference r is
    hash_to_g2(transcript) (ChaChaRng); the boundary takes it as an input."""
    s_delta = G1.mul(s, delta)
    h = hashlib.blake2b()
    h.update(d["cs_hash"])
    for pk in d["contributions"]:
        h.update(pk)
    h.update(g1_encode(s, False))
    h.update(g1_encode(s_delta, False))
    transcript = h.digest()
    r_delta = G2.mul(r_g2, delta)
    delta_after = G1.mul(g1_decode(d["delta_g1"], False), delta)
    return (g1_encode(delta_after, False) + g1_encode(s, False) + g1_encode(s_delta, False) +
            g2_encode(r_delta, False) + transcript)


def phase2_contribute(buf, delta, s, r_g2):
    """phase2/src/parameters.rs:414-522: returns (new params bytes, 64-byte contribution hash)."""
    d = params_parse(buf)
    pubkey = phase2_pubkey(d, delta, s, r_g2)
    dinv = pow(delta, -1, R)
    d["l"] = [g1_encode(G1.mul(g1_decode(e, False), dinv), False) for e in d["l"]]
    d["h"] = [g1_encode(G1.mul(g1_decode(e, False), dinv), False) for e in d["h"]]
    d["delta_g1"] = g1_encode(G1.mul(g1_decode(d["delta_g1"], False), delta), False)
    d["delta_g2"] = g2_encode(G2.mul(g2_decode(d["delta_g2"], False), delta), False)
    d["contributions"] = d["contributions"] + [pubkey]
    return params_serialize(d), hashlib.blake2b(pubkey).digest()


# ----------------------------------------------------------------------------- MSM / FFT
def msm(curve, points, scalars):
    acc = None
    for p, k in zip(points, scalars):
        acc = curve.add(acc, curve.mul(p, k % R))
    return acc


def domain_omega(log_n):
    """bellman/src/domain.rs:80-89."""
    w = FR_ROOT_OF_UNITY
    for _ in range(log_n, FR_S):
        w = w * w % R
    return w


def fft(a, inverse=False, coset=False):
    """Natural-order in / natural-order out DFT of length 2^k (domain.rs:154-195).
    coset fft: scale by g^i first; coset ifft: ifft then scale by g^-i."""
    n = len(a)
    log_n = n.bit_length() - 1
    assert 1 << log_n == n
    w = domain_omega(log_n)
    a = list(a)
    if coset and not inverse:
        g = 1
        for i in range(n):
            a[i] = a[i] * g % R
            g = g * FR_GENERATOR % R
    if inverse:
        w = pow(w, -1, R)

    def rec(v, w):
        if len(v) == 1:
            return v
        w2 = w * w % R
        ev, od = rec(v[0::2], w2), rec(v[1::2], w2)
        h = len(v) // 2
        out = [0] * len(v)
        t = 1
        for i in range(h):
            x = od[i] * t % R
            out[i] = (ev[i] + x) % R
            out[i + h] = (ev[i] - x) % R
            t = t * w % R
        return out

    a = rec(a, w)
    if inverse:
        ninv = pow(n, -1, R)
        a = [x * ninv % R for x in a]
        if coset:
            gi = pow(FR_GENERATOR, -1, R)
            g = 1
            for i in range(n):
                a[i] = a[i] * g % R
                g = g * gi % R
    return a


# ----------------------------------------------------------------------------- GLV constants
def glv_constants():
    """beta in Fq, lambda in Fr with phi(x,y) = (beta x, y) = [lambda](x,y) on G1."""
    beta = next(b for b in (pow(g, (Q - 1) // 3, Q) for g in range(2, 50)) if b != 1)
    lam = next(l for l in (pow(g, (R - 1) // 3, R) for g in range(2, 50)) if l != 1)
    for lm in (lam, lam * lam % R):
        if G1.mul(G1_GEN, lm) == (beta * G1_GEN[0] % Q, G1_GEN[1]):
            return beta, lm
    raise AssertionError("no matching lambda")


if __name__ == "__main__":
    b, l = glv_constants()
    print("beta  =", hex(b))
    print("lambda=", hex(l))
