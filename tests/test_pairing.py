"""CPU: the verifier's host side inside libp2b.so (csrc/pairing.cuh, host_verify.cu) -- pairing checks, hash_to_g2, the
ChaCha key-generation RNG and the phase-1 / phase-2 keypair mirrors.  No GPU needed: this code runs on the host in the
reference as well (powersoftau/src/utils.rs:151-159, phase2/src/utils.rs:48-57).

What pins it: bilinearity and non-degeneracy against the oracle's scalar multiplications (a bilinear, non-degenerate map
decides same_ratio uniquely), the reference's hard-coded Frobenius tables (pairing/src/bn256/fq.rs:106-119,121-199,
280-432) for the constants computed at start-up, and the published ChaCha20 keystream for the RNG core.  hash_to_g2 and
the samplers follow rand 0.4.6 / ff_derive, which are not vendored under the reference tree: their byte-level output has
no reference vector in-tree (parity unpinned); the tests check the structural properties."""
import hashlib
import random

import numpy as np
import pytest

from util import G1_GEN, G2_GEN, Q_MOD, R_MOD, be


@pytest.fixture(scope="module")
def lib():
    from phase2_bn254_b200 import lib as L
    L.load()
    return L


def neg_g1(p):
    return p[:32] + be((-int.from_bytes(p[32:], "big")) % Q_MOD)


def test_same_ratio_reference_cases(lib, oracle):
    """test_same_ratio_bn256 of powersoftau/src/utils.rs:76-88."""
    rng = random.Random(1)
    for _ in range(3):
        s = rng.randrange(1, R_MOD)
        g1_s, g2_s = oracle.point_mul(0, G1_GEN, be(s)), oracle.point_mul(1, G2_GEN, be(s))
        assert lib.same_ratio((G1_GEN, g1_s), (G2_GEN, g2_s))
        assert not lib.same_ratio((g1_s, G1_GEN), (G2_GEN, g2_s))
    zero1, zero2 = bytes([0x40]) + bytes(63), bytes([0x40]) + bytes(127)
    assert not lib.same_ratio((zero1, zero1), (G2_GEN, G2_GEN))          # utils.rs:155-157: any zero => false
    assert not lib.same_ratio((G1_GEN, G1_GEN), (zero2, G2_GEN))
    assert lib.same_ratio((G1_GEN, G1_GEN), (G2_GEN, G2_GEN))


def test_bilinearity_and_non_degeneracy(lib, oracle):
    """random_bilinearity_tests of pairing/src/bn256/mod.rs:559-600, stated on products: e(aP, bQ) e(-abP, Q) = 1."""
    rng = random.Random(2)
    assert not lib.pairing_check(G1_GEN, G2_GEN)                          # e(G, H) != 1
    assert lib.pairing_check(b"", b"")
    for _ in range(3):
        a, b, c = (rng.randrange(1, R_MOD) for _ in range(3))
        p = oracle.point_mul(0, G1_GEN, be(c))
        q = oracle.point_mul(1, G2_GEN, be(c * 7 % R_MOD))
        ap, bq = oracle.point_mul(0, p, be(a)), oracle.point_mul(1, q, be(b))
        abp = oracle.point_mul(0, p, be(a * b % R_MOD))
        assert lib.pairing_check(ap + neg_g1(abp), bq + q)
        assert not lib.pairing_check(ap + neg_g1(abp), q + bq)
        # e(aP, Q) = e(P, aQ)
        assert lib.same_ratio((p, ap), (q, oracle.point_mul(1, q, be(a))))
    # infinity pairs contribute the identity
    zero1 = bytes([0x40]) + bytes(63)
    assert lib.pairing_check(zero1, G2_GEN)


def test_decode_errors(lib):
    bad = G1_GEN[:63] + b"\x03"                                           # (1, 3) is not on the curve
    with pytest.raises(lib.P2BError) as e:
        lib.same_ratio((bad, G1_GEN), (G2_GEN, G2_GEN))
    assert e.value.code == lib.EDECODE


def test_frobenius_constants_match_reference_tables(lib):
    """gamma = xi^((q-1)/6), gamma^2, gamma^3 computed at start-up == FROBENIUS_COEFF_FQ12_C1[1],
    FROBENIUS_COEFF_FQ6_C1[1], XI_TO_Q_MINUS_1_OVER_2 (Montgomery limbs as written in pairing/src/bn256/fq.rs)."""
    import ctypes
    out = (ctypes.c_uint8 * 192)()
    assert lib.load().p2b_pairing_constants(out) == 0
    raw = bytes(out)

    def limbs(*v):
        return b"".join(int(x).to_bytes(8, "little") for x in v)

    fq12_c1_1 = (limbs(0xaf9ba69633144907, 0xca6b1d7387afb78a, 0x11bded5ef08a2087, 0x02f34d751a1f3a7c) +
                 limbs(0xa222ae234c492d72, 0xd00f02a4565de15b, 0xdc2ff3a253dfc926, 0x10a75716b3899551))   # fq.rs:291-304
    fq6_c1_1 = (limbs(0xb5773b104563ab30, 0x347f91c8a9aa6454, 0x7a007127242e0991, 0x1956bcd8118214ec) +
                limbs(0x6e849f1ea0aa4757, 0xaa1c7b6d89f89141, 0xb6e713cdfae0ca3a, 0x26694fbb4e82ebc3))    # fq.rs:133-147
    xi_qm1_2 = (limbs(0xe4bbdd0c2936b629, 0xbb30f162e133bacb, 0x31a9d1b6f9645366, 0x253570bea500f8dd) +
                limbs(0xa1d77ce45ffe77c7, 0x07affd117826d1db, 0x6d16bd27bb7edc6b, 0x2c87200285defecc))    # fq.rs:106-119
    assert raw[:64] == fq12_c1_1
    assert raw[64:128] == fq6_c1_1
    assert raw[128:] == xi_qm1_2


def _chacha_block(key_words, counter):
    """Independent restatement of the ChaCha20 block function (128-bit counter in words 12..15)."""
    st = [0x61707865, 0x3320646e, 0x79622d32, 0x6b206574] + list(key_words) + [(counter >> (32 * i)) & 0xffffffff for i in range(4)]
    x = list(st)
    rot = lambda v, c: ((v << c) | (v >> (32 - c))) & 0xffffffff

    def qr(a, b, c, d):
        x[a] = (x[a] + x[b]) & 0xffffffff; x[d] = rot(x[d] ^ x[a], 16)
        x[c] = (x[c] + x[d]) & 0xffffffff; x[b] = rot(x[b] ^ x[c], 12)
        x[a] = (x[a] + x[b]) & 0xffffffff; x[d] = rot(x[d] ^ x[a], 8)
        x[c] = (x[c] + x[d]) & 0xffffffff; x[b] = rot(x[b] ^ x[c], 7)

    for _ in range(10):
        qr(0, 4, 8, 12); qr(1, 5, 9, 13); qr(2, 6, 10, 14); qr(3, 7, 11, 15)
        qr(0, 5, 10, 15); qr(1, 6, 11, 12); qr(2, 7, 8, 13); qr(3, 4, 9, 14)
    return [(a + b) & 0xffffffff for a, b in zip(x, st)]


def test_chacha_keystream(lib):
    """Zero key: the published ChaCha20 keystream (first block 76 b8 e0 ad a0 f1 3d 90 ..., the vector of rand 0.4's
    test_rng_true_values); then random seeds across several blocks against the independent restatement above."""
    rng = lib.ChaChaRng([0] * 8)
    first = [rng.next_u32() for _ in range(16)]
    assert first[:4] == [0xade0b876, 0x903df1a0, 0xe56a5d40, 0x28bd8653]
    assert b"".join(w.to_bytes(4, "little") for w in first).hex().startswith("76b8e0ada0f13d90405d6ae55386bd28bdd219b8a08ded1aa836efcc8b770dc7")
    assert first == _chacha_block([0] * 8, 0)
    assert [rng.next_u32() for _ in range(16)] == _chacha_block([0] * 8, 1)
    r = random.Random(3)
    seed = [r.getrandbits(32) for _ in range(8)]
    rng = lib.ChaChaRng(seed)
    got = [rng.next_u32() for _ in range(16 * 5 + 3)]
    exp = sum((_chacha_block(seed, c) for c in range(6)), [])
    assert got == exp[: len(got)]


def _mont_inv():
    return pow(1 << 256, -1, R_MOD), pow(1 << 256, -1, Q_MOD)


def test_samplers_follow_the_documented_recipe(lib, oracle):
    """Fr::rand / G1::rand / G2::rand over the ChaCha stream, re-derived here with big integers from the raw words:
    four u64 limbs (each (next_u32 << 32) | next_u32), two top bits shaved, rejection, the limbs being the MONTGOMERY
    form; G1: x, then `greatest` = next_u32 & 1, y = the root picked by ((y < -y) ^ greatest)."""
    seed = [7, 6, 5, 4, 3, 2, 1, 0]
    words = sum((_chacha_block(seed, c) for c in range(8)), [])
    pos = 0

    def u64():
        nonlocal pos
        v = (words[pos] << 32) | words[pos + 1]
        pos += 2
        return v

    def field(mod):
        while True:
            limbs = [u64() for _ in range(4)]
            limbs[3] &= (1 << 62) - 1
            v = sum(l << (64 * i) for i, l in enumerate(limbs))
            if v < mod:
                return v

    rinv, qinv = _mont_inv()
    rng = lib.ChaChaRng(seed)
    for _ in range(3):
        assert rng.gen_fr() == field(R_MOD) * rinv % R_MOD
    for _ in range(3):
        got = rng.gen_g1()
        while True:
            x = field(Q_MOD) * qinv % Q_MOD
            greatest = words[pos] & 1
            pos += 1
            y2 = (x * x * x + 3) % Q_MOD
            y = pow(y2, (Q_MOD + 1) // 4, Q_MOD)
            if y * y % Q_MOD == y2:
                break
        y = max(y, Q_MOD - y) if greatest else min(y, Q_MOD - y)
        assert got == be(x) + be(y)


def test_hash_to_g2_properties(lib, oracle):
    """test_hash_to_g2_bn256 of powersoftau/src/utils.rs:50-74 (only the first 32 bytes matter) + the result is a point
    of order r on the twist."""
    d = bytes(range(1, 34))
    a = lib.hash_to_g2(d)
    assert a == lib.hash_to_g2(d[:32] + b"\x22")
    assert a != lib.hash_to_g2(d[:31] + b"\x21")
    assert a == lib.ChaChaRng.from_digest(d).gen_g2()
    assert oracle.point_recode(1, a, 0, 0, checked=True) == a            # on the curve
    minus = oracle.point_mul(1, a, be(R_MOD - 1))                         # [r - 1]P == -P  <=>  [r]P == 0
    y1, y0 = int.from_bytes(a[64:96], "big"), int.from_bytes(a[96:], "big")
    assert minus == a[:64] + be(-y1 % Q_MOD) + be(-y0 % Q_MOD)


def test_host_mul_vs_oracle(lib, oracle):
    rng = random.Random(4)
    for k in (0, 1, 2, R_MOD - 1, rng.randrange(R_MOD), rng.randrange(R_MOD)):
        assert lib.host_mul(0, G1_GEN, be(k)) == oracle.point_mul(0, G1_GEN, be(k))
        assert lib.host_mul(1, G2_GEN, be(k)) == oracle.point_mul(1, G2_GEN, be(k))
    with pytest.raises(lib.P2BError):
        lib.host_mul(0, G1_GEN, be(R_MOD))                               # scalars are canonical


def test_phase1_keypair_verifies(lib):
    """keypair(rng, digest) (keypair.rs:54-103): the three proofs of knowledge hold, (de)serialisation round-trips."""
    from phase2_bn254_b200.powersoftau import PublicKey, compute_g2_s, keypair, same_ratio, DeserializationError
    digest = hashlib.blake2b(b"transcript").digest()
    pub, priv = keypair(lib.ChaChaRng.from_digest(hashlib.sha256(b"seed").digest()), digest)
    assert 0 < priv.tau < R_MOD and 0 < priv.alpha < R_MOD and 0 < priv.beta < R_MOD
    for pers, g1, g2 in ((0, pub.tau_g1, pub.tau_g2), (1, pub.alpha_g1, pub.alpha_g2), (2, pub.beta_g1, pub.beta_g2)):
        g2_s = compute_g2_s(digest, g1[0], g1[1], pers)
        assert same_ratio(g1, (g2_s, g2))
        assert not same_ratio(g1, (g2, g2_s))
    raw = pub.serialize()
    assert len(raw) == 768 and PublicKey.deserialize(raw) == pub
    with pytest.raises(DeserializationError):
        PublicKey.deserialize(raw[:63] + bytes([raw[63] ^ 1]) + raw[64:])
    with pytest.raises(DeserializationError):
        PublicKey.deserialize(bytes([0x40]) + bytes(63) + raw[64:])
    # same seed, same key (the beacon binaries rely on this)
    pub2, priv2 = keypair(lib.ChaChaRng.from_digest(hashlib.sha256(b"seed").digest()), digest)
    assert pub2 == pub and priv2 == priv


def test_keygen_regression_vectors(lib):
    """tests/golden/keygen_selfgen.json: SELF-GENERATED vectors (tools/make_keygen_golden.py) -- they pin the host RNG path
    against accidental change; they are not reference output (no reference vector exists for this path in-tree)."""
    import json
    import os
    from phase2_bn254_b200.powersoftau import keypair
    g = json.load(open(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "keygen_selfgen.json")))
    rng = lib.ChaChaRng(g["seed"])
    assert [hex(rng.gen_fr()) for _ in range(3)] == g["fr"]
    assert rng.gen_g1().hex() == g["g1"] and rng.gen_g2().hex() == g["g2"] and rng.next_u32() == g["u32_after"]
    assert lib.hash_to_g2(bytes(range(32))).hex() == g["hash_to_g2_of_0_to_31"]
    k = g["phase1_keypair"]
    pub, priv = keypair(lib.ChaChaRng.from_digest(hashlib.sha256(b"golden seed").digest()), bytes.fromhex(k["digest"]))
    assert (hex(priv.tau), hex(priv.alpha), hex(priv.beta)) == (k["tau"], k["alpha"], k["beta"])
    assert hashlib.blake2b(pub.serialize()).hexdigest() == k["public_key_blake2b"]
