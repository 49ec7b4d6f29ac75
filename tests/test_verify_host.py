"""CPU: the host logic of the verifier / key-generation / MPCParameters.new mirrors (phase2_bn254_b200/{powersoftau,phase2}.py)
with the oracle standing in for the GPU calls (MSM, bulk codec): chunk walk and ratio checks of verify_transformation,
the structural and signature checks of verify_contribution, KeypairAssembly / CSR flattening.  The same flows run on the
real GPU path in tests/test_gpu_verify.py and tests/test_gpu_phase2_new.py."""
import hashlib
import struct

import numpy as np
import pytest

from util import G1_GEN, G2_GEN, R_MOD, be


from util import OracleCtx as _OracleCtx


class OracleCtx(_OracleCtx):
    pass


@pytest.fixture(scope="module")
def lib():
    from phase2_bn254_b200 import lib as L
    L.load()
    return L


def test_verify_transformation_host_logic(lib, oracle):
    from phase2_bn254_b200.powersoftau import (CeremonyParams, PrivateKey, PublicKey, calculate_hash, public_key_for,
                                               verify_transformation)
    size, batch = 3, 4
    params = CeremonyParams(size, batch)
    ctx = OracleCtx(oracle)
    ch0 = np.frombuffer(oracle.pot_generate_initial(size), dtype=np.uint8)
    digest = calculate_hash(ch0)
    key = PrivateKey(0x1234567 % R_MOD, 0x89abcdef % R_MOD, 0x13579bdf % R_MOD)
    pub = public_key_for(key, lib.ChaChaRng([5] * 8), digest)
    body = oracle.pot_transform(ch0.tobytes(), size, batch, be(key.tau), be(key.alpha), be(key.beta), threads=4)
    rs = np.zeros(params.contribution_size, dtype=np.uint8)
    rs[:len(body)] = np.frombuffer(body, dtype=np.uint8)
    rs[:64] = np.frombuffer(digest, dtype=np.uint8)
    pub.write(rs, True, params)
    assert PublicKey.read(rs, True, params) == pub
    rng = np.random.default_rng(1)
    assert verify_transformation(ch0, rs, pub, digest, False, True, True, True, params, ctx=ctx, rng=rng)
    # a different chunking of the same walk (the verifier's batch size is its own choice), 128-bit random coefficients
    assert verify_transformation(ch0, rs, pub, digest, False, True, False, True, CeremonyParams(size, 8), ctx=ctx, rng=rng,
                                 scalar_bits=128)
    # wrong digest, key of other secrets, tampered elements in both loops and at the intersection
    assert not verify_transformation(ch0, rs, pub, hashlib.blake2b(b"x").digest(), False, True, False, True, params, ctx=ctx, rng=rng)
    other = public_key_for(PrivateKey(key.tau + 1, key.alpha, key.beta), lib.ChaChaRng([5] * 8), digest)
    assert not verify_transformation(ch0, rs, other, digest, False, True, False, True, params, ctx=ctx, rng=rng)
    foreign = rs[64 + 32 * 2: 64 + 32 * 3].copy()
    for idx in (5, params.powers_length, params.powers_g1_length - 1):
        bad = rs.copy()
        bad[64 + 32 * idx: 64 + 32 * (idx + 1)] = foreign
        assert not verify_transformation(ch0, bad, pub, digest, False, True, False, True, params, ctx=ctx, rng=rng), idx
    with pytest.raises(RuntimeError):                                    # the reference panics on a one-element chunk
        verify_transformation(ch0, rs, pub, digest, False, True, False, True, CeremonyParams(size, 1), ctx=ctx, rng=rng)


def _params(lib, m, delta_g1=G1_GEN, delta_g2=G2_GEN):
    g1 = lambda i: lib.host_mul(0, G1_GEN, be(1000 + i))
    g2 = lambda i: lib.host_mul(1, G2_GEN, be(2000 + i))
    body = g1(1) + g1(2) + g2(3) + g2(4) + delta_g1 + delta_g2
    for n, s, grp in ((2, 10, 0), (m - 1, 20, 0), (m, 40, 0), (2, 60, 0), (2, 70, 0), (2, 80, 1)):
        body += struct.pack(">I", n) + b"".join((g2 if grp else g1)(s + i) for i in range(n))
    body += hashlib.blake2b(body).digest()
    return body + struct.pack(">I", 0)


def contribution_by_hand(lib, m):
    """(before, after bytes, contribution hash, layout): keypair() + host scalar multiplications, no GPU."""
    from phase2_bn254_b200.phase2 import MPCParameters, keypair, params_layout
    before = MPCParameters(_params(lib, m))
    pk, delta = keypair(lib.ChaChaRng([8, 7, 6, 5, 4, 3, 2, 1]), before)
    assert len(pk) == 384 and 0 < delta < R_MOD
    lay = params_layout(before.data)
    dinv = be(pow(delta, -1, R_MOD))
    after = bytearray(before.data.tobytes())
    for name in ("h", "l"):
        off, n, size, _ = lay[name]
        for i in range(n):
            after[off + 64 * i: off + 64 * (i + 1)] = lib.host_mul(0, bytes(after[off + 64 * i: off + 64 * (i + 1)]), dinv)
    o1, o2 = lay["delta_g1"][0], lay["delta_g2"][0]
    after[o1:o1 + 64] = lib.host_mul(0, bytes(after[o1:o1 + 64]), be(delta))
    after[o2:o2 + 128] = lib.host_mul(1, bytes(after[o2:o2 + 128]), be(delta))
    after = bytes(after[:-4]) + struct.pack(">I", 1) + pk
    return before, after, hashlib.blake2b(pk).digest(), lay


def test_verify_contribution_host_logic(lib, oracle):
    """A contribution assembled by hand on the host (keypair + host scalar multiplications), verified with the oracle's MSM."""
    from phase2_bn254_b200.phase2 import MPCParameters, VerificationError, verify_contribution
    ctx = OracleCtx(oracle)
    before, after, expected, lay = contribution_by_hand(lib, 4)
    pk = after[-384:]
    o1, o2 = lay["delta_g1"][0], lay["delta_g2"][0]
    rng = np.random.default_rng(2)
    assert verify_contribution(before, MPCParameters(after), ctx=ctx, rng=rng) == hashlib.blake2b(pk).digest()
    assert verify_contribution(before, MPCParameters(after), ctx=ctx, rng=rng, scalar_bits=128) == hashlib.blake2b(pk).digest()
    # structural rejections need no MSM at all (ctx=None: they must fail before the GPU is asked for)
    with pytest.raises(VerificationError):
        verify_contribution(before, before, ctx=None, rng=rng)
    bad = bytearray(after); bad[lay["a"][0] + 5] ^= 1
    with pytest.raises(VerificationError):
        verify_contribution(before, MPCParameters(bytes(bad)), ctx=None, rng=rng)
    bad = bytearray(after); bad[-1] ^= 1                                 # transcript
    with pytest.raises(VerificationError):
        verify_contribution(before, MPCParameters(bytes(bad)), ctx=None, rng=rng)
    bad = bytearray(after); bad[o2:o2 + 128] = before.data[o2:o2 + 128].tobytes()      # delta_g2 not updated
    with pytest.raises(VerificationError):
        verify_contribution(before, MPCParameters(bytes(bad)), ctx=None, rng=rng)
    off = lay["h"][0]
    bad = bytearray(after); bad[off:off + 64] = before.data[off:off + 64].tobytes()    # one H element not updated
    with pytest.raises(VerificationError):
        verify_contribution(before, MPCParameters(bytes(bad)), ctx=ctx, rng=rng)


def test_keypair_assembly_and_csr():
    """keypair_assembly.rs:20-118: every term of the three linear combinations lands in its variable's row with the index of
    the constraint; coefficients are reduced mod r; rows flatten to CSR in input-then-aux order."""
    from phase2_bn254_b200.phase2 import KeypairAssembly, _csr
    cs = KeypairAssembly()
    one = cs.alloc_input()
    x = cs.alloc_input()
    a, b = cs.alloc(), cs.alloc()
    assert (one, x, a, b) == (("input", 0), ("input", 1), ("aux", 0), ("aux", 1))
    cs.enforce([(a, 1), (one, -1)], [(b, 3)], [(x, 1)])
    cs.enforce([(a, 2)], [(a, 2)], [(b, R_MOD + 5)])
    assert cs.num_constraints == 2 and cs.num_inputs == 2 and cs.num_aux == 2
    assert cs.at_inputs == [[(R_MOD - 1, 0)], []] and cs.at_aux == [[(1, 0), (2, 1)], []]
    assert cs.bt_aux == [[(2, 1)], [(3, 0)]] and cs.ct_inputs == [[], [(1, 0)]] and cs.ct_aux == [[], [(5, 1)]]
    offs, cols, k = _csr(cs.at_inputs + cs.at_aux, col_shift=7)
    assert offs.tolist() == [0, 1, 1, 3, 3] and cols.tolist() == [7, 7, 8]
    assert k.tobytes() == be(R_MOD - 1) + be(1) + be(2)


def test_random_scalars_shape():
    from phase2_bn254_b200.powersoftau import _random_scalars
    rng = np.random.default_rng(3)
    a = _random_scalars(rng, 200, 253).reshape(200, 32)
    assert a[:, 0].max() <= 0x1f and a[:, 0].max() > 0 and all(int.from_bytes(r.tobytes(), "big") < R_MOD for r in a)
    b = _random_scalars(rng, 200, 128).reshape(200, 32)
    assert not b[:, :16].any() and b[:, 16:].any()
    with pytest.raises(ValueError):
        _random_scalars(rng, 1, 254)


def test_mpc_parameters_new_host_logic(lib, oracle):
    """MPCParameters.new orchestration (radix-file parsing, ONE input + x*0=0 constraints, the ext = A.beta + B.alpha + C.coeffs
    concatenation, filtering, cs_hash, layout) with the oracle's eval restatement standing in for p2b_g{1,2}_sparse_mul."""
    from phase2_bn254_b200.phase2 import KeypairAssembly, MPCParameters, SynthesisError, params_layout

    class Ctx(OracleCtx):
        def sparse_mul(self, group, bases, row_offsets, cols, coeffs):
            return np.frombuffer(self.oc.sparse_mul(group, bytes(np.asarray(bases)), list(row_offsets), list(cols),
                                                    bytes(np.asarray(coeffs))), dtype=np.uint8)

    m = 8
    g1 = lambda i: lib.host_mul(0, G1_GEN, be(500 + i))
    g2 = lambda i: lib.host_mul(1, G2_GEN, be(700 + i))
    coeffs_g1 = [g1(i) for i in range(m)]
    coeffs_g2 = [g2(i) for i in range(m)]
    alpha_c = [g1(100 + i) for i in range(m)]
    beta_c = [g1(200 + i) for i in range(m)]
    h = [g1(300 + i) for i in range(m - 1)]
    radix = g1(1) + g1(2) + g2(3) + b"".join(coeffs_g1) + b"".join(coeffs_g2) + b"".join(alpha_c) + b"".join(beta_c) + b"".join(h)
    assert len(radix) == 192 + 384 * m

    def circuit(cs):                      # x * y = z ; (z + 1) * 1 = out ; y is used only in B
        out = cs.alloc_input()
        x, y, z = cs.alloc(), cs.alloc(), cs.alloc()
        one = ("input", 0)
        cs.enforce([(x, 1)], [(y, 1)], [(z, 1)])
        cs.enforce([(z, 1), (one, 1)], [(one, 1)], [(out, 1)])
        cs.enforce([(x, 3)], [(one, R_MOD - 2)], [(x, R_MOD - 6)])

    ctx = Ctx(oracle)
    seen = []
    p = MPCParameters.new(circuit, False, lambda exp: (seen.append(exp), radix)[1], ctx=ctx)
    assert seen == [3]                                                   # 3 + 2 input constraints = 5 -> m = 8
    lay = params_layout(p.data)
    assert (lay["ic"][1], lay["l"][1], lay["a"][1], lay["b_g1"][1], lay["b_g2"][1], lay["h"][1]) == (2, 3, 5, 5, 5, m - 1)
    sec = lambda name: p.section(name).tobytes()
    assert sec("alpha_g1") == g1(1) and sec("beta_g1") == g1(2) and sec("beta_g2") == g2(3) and sec("h") == b"".join(h)
    assert sec("delta_g1") == G1_GEN and sec("gamma_g2") == G2_GEN and sec("delta_g2") == G2_GEN
    pm = lambda pt, k: oracle.point_mul(0, pt, be(k % R_MOD))
    add = lambda *pts: oracle.sum_points(0, b"".join(pts))
    # variable order: ONE, out, x, y, z.  A rows: ONE {c1:1, c3(input constraint 0):1}, out {c4:1}, x {c0:1, c2:3}, y {}, z {c1:1}
    a = sec("a")
    assert a[0:64] == add(coeffs_g1[1], coeffs_g1[3]) and a[64:128] == coeffs_g1[4]
    assert a[128:192] == add(coeffs_g1[0], pm(coeffs_g1[2], 3))
    assert a[192:256] == bytes([0x40]) + bytes(63) and a[256:320] == coeffs_g1[1]
    # ext for x (an aux variable -> l[0]): A.beta {c0:1, c2:3} + B.alpha {} + C.coeffs {c2:-6}
    assert sec("l")[:64] == add(beta_c[0], pm(beta_c[2], 3), pm(coeffs_g1[2], -6))
    # ext for out (an input -> ic[1]): A.beta {c4:1} + C.coeffs {c1:1}
    assert sec("ic")[64:128] == add(beta_c[4], coeffs_g1[1])
    assert sec("cs_hash") == hashlib.blake2b(p.data[:lay["cs_hash"][0]].tobytes()).digest()
    # filtering drops the infinity of y out of A; in B only ONE (constraints 1, 2) and y (constraint 0) appear
    pf = MPCParameters.new(circuit, True, lambda exp: radix, ctx=ctx)
    layf = params_layout(pf.data)
    assert (layf["a"][1], layf["b_g1"][1], layf["b_g2"][1]) == (4, 2, 2)
    assert pf.section("b_g1").tobytes() == add(coeffs_g1[1], pm(coeffs_g1[2], -2)) + coeffs_g1[0]
    assert pf.section("b_g2").tobytes()[128:] == coeffs_g2[0]
    # an aux variable that appears nowhere
    def loose(cs):
        circuit(cs)
        cs.alloc()
    with pytest.raises(SynthesisError):
        MPCParameters.new(loose, False, lambda exp: radix, ctx=ctx)
    with pytest.raises(IOError):
        MPCParameters.new(circuit, False, lambda exp: radix[:-1], ctx=ctx)


def test_mpc_parameters_read_is_as_strict_as_the_reference(lib, oracle):
    """MPCParameters::read: the verifying key and the stored contributions are validated whatever the caller's flags say
    (VerifyingKey::read groth16/mod.rs:160-199 always uses into_affine() and rejects infinity in ic; PublicKey::read
    phase2/src/keypair.rs:64-100 rejects infinity and off-curve points); only h, l, a, b_g1, b_g2 follow `checked` /
    `disallow_points_at_infinity`."""
    from phase2_bn254_b200.phase2 import MPCParameters, params_layout
    ctx = OracleCtx(oracle)
    _, after, _, _ = contribution_by_hand(lib, 4)
    lay = params_layout(after)
    assert MPCParameters.read(after, ctx=ctx).validated
    assert not MPCParameters.read(after, checked=False, ctx=ctx).validated
    inf1, inf2 = bytes([0x40]) + bytes(63), bytes([0x40]) + bytes(127)

    def tampered(off, repl):
        b = bytearray(after)
        b[off: off + len(repl)] = repl
        return bytes(b)

    off_curve = bytearray(after[lay["alpha_g1"][0]: lay["alpha_g1"][0] + 64])
    off_curve[63] ^= 1
    cases = [("alpha_g1 off curve", tampered(lay["alpha_g1"][0], bytes(off_curve))),
             ("ic infinity", tampered(lay["ic"][0] + 64, inf1)),
             ("contribution s infinity", tampered(lay["contributions"][0] + 64, inf1)),
             ("contribution s_delta off curve", tampered(lay["contributions"][0] + 128, bytes(off_curve))),
             ("contribution r_delta infinity", tampered(lay["contributions"][0] + 192, inf2))]
    for what, buf in cases:
        for checked in (False, True):
            with pytest.raises(IOError):
                MPCParameters.read(buf, disallow_points_at_infinity=False, checked=checked, ctx=ctx)
    # the caller's flags govern the query vectors: infinity in h is tolerated unless disallowed, off-curve only seen when checked
    h_inf = tampered(lay["h"][0], inf1)
    MPCParameters.read(h_inf, ctx=ctx)
    with pytest.raises(IOError):
        MPCParameters.read(h_inf, disallow_points_at_infinity=True, ctx=ctx)
    h_off = tampered(lay["h"][0], bytes(off_curve))
    MPCParameters.read(h_off, checked=False, ctx=ctx)
    with pytest.raises(IOError):
        MPCParameters.read(h_off, checked=True, ctx=ctx)
    # vk infinity is not rejected by VerifyingKey::read for the single elements (only ic)
    MPCParameters.read(tampered(lay["gamma_g2"][0], inf2), ctx=ctx)


def test_verify_contribution_revalidates_unchecked_parameters(lib, oracle):
    """The MSM decodes unchecked; an MPCParameters built from raw bytes (no read(checked=True)) has its H / L put through the
    checked codec by verify_contribution, so an off-curve element is a VerificationError, not a silently wrong sum."""
    from phase2_bn254_b200.phase2 import MPCParameters, VerificationError, verify_contribution
    ctx = OracleCtx(oracle)
    before, after, expected, lay = contribution_by_hand(lib, 4)
    assert verify_contribution(before, MPCParameters(after), ctx=ctx, rng=np.random.default_rng(1)) == expected
    bad = bytearray(after)
    bad[lay["l"][0] + 64 + 63] ^= 1
    with pytest.raises(VerificationError):
        verify_contribution(before, MPCParameters(bytes(bad)), ctx=ctx, rng=np.random.default_rng(1))


def test_scalar_bits_floor_and_system_rng():
    from phase2_bn254_b200.powersoftau import _random_scalars, system_rng
    for bits in (8, 64, 120):
        with pytest.raises(ValueError):
            _random_scalars(system_rng(), 4, bits)
    a, b = _random_scalars(system_rng(), 64, 128), _random_scalars(system_rng(), 64, 128)
    assert a.size == 64 * 32 and not np.array_equal(a, b)
    assert not a.reshape(64, 32)[:, :16].any() and a.reshape(64, 32)[:, 16:].any()
    assert (_random_scalars(system_rng(), 64, 253).reshape(64, 32)[:, 0] < 0x20).all()


def test_contribute_challenge_overlapped_equals_reference_order(lib, oracle):
    """`contribute_challenge` (the body of compute_constrained.rs:140-230) with the challenge hashed on a host thread next to the
    transform writes the same response as the reference's order of operations, consumes the generator in keypair()'s order
    (keypair.rs:54-103), and its output passes verify_transformation."""
    from phase2_bn254_b200.powersoftau import (CeremonyParams, PublicKey, calculate_hash, contribute_challenge, keypair,
                                               verify_transformation)
    size, batch = 3, 4
    params = CeremonyParams(size, batch)

    class Ctx(OracleCtx):
        def pot_accumulator_size(self, size, compressed):
            p = CeremonyParams(size, 4)
            return p.contribution_size - p.public_key_size if compressed else p.accumulator_size

        def pot_transform(self, imap, omap, size, batch, tau, alpha, beta, in_c, out_c, check, shard_index=0, shard_count=1):
            body = self.oc.pot_transform(bytes(np.asarray(imap)), size, batch, bytes(tau), bytes(alpha), bytes(beta), in_c, out_c,
                                         bool(check & 1), threads=4)
            omap[64:len(body)] = np.frombuffer(body, dtype=np.uint8)[64:]

    ctx = Ctx(oracle)
    ch = np.frombuffer(oracle.pot_generate_initial(size), dtype=np.uint8)
    out = {}
    for overlap in (True, False):
        rs = np.zeros(params.contribution_size, dtype=np.uint8)
        d, h, pub = contribute_challenge(ch, rs, lib.ChaChaRng([7] * 8), params, ctx=ctx, overlap=overlap)
        assert d == calculate_hash(ch) == bytes(rs[:64]) and h == calculate_hash(rs)
        out[overlap] = (rs.tobytes(), h, pub)
    assert out[True] == out[False]
    rs = np.frombuffer(out[True][0], dtype=np.uint8)
    pub_ref, key_ref = keypair(lib.ChaChaRng([7] * 8), calculate_hash(ch))          # the reference's single call
    assert pub_ref == out[True][2] == PublicKey.read(rs, True, params)
    body = oracle.pot_transform(ch.tobytes(), size, batch, be(key_ref.tau), be(key_ref.alpha), be(key_ref.beta), threads=4)
    assert rs[64:len(body)].tobytes() == body[64:]
    assert verify_transformation(ch, rs, pub_ref, calculate_hash(ch), False, True, True, True, params, ctx=ctx,
                                 rng=np.random.default_rng(2))
