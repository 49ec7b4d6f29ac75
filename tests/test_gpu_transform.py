"""GPU parity: powersoftau transform and phase2 contribute on the reference's file formats vs the oracle."""
import hashlib
import struct

import numpy as np
import pytest

from util import G1_GEN, G2_GEN, R_MOD, be, random_points

pytestmark = pytest.mark.gpu

TAU = 0x1111111111111111111111111111111111111111111111111111111111111111 % R_MOD
ALPHA = 0x2222222222222222222222222222222222222222222222222222222222222222 % R_MOD
BETA = 0x0333333333333333333333333333333333333333333333333333333333333333 % R_MOD


@pytest.mark.parametrize("size,batch", [(3, 4), (6, 16), (10, 256)])
def test_transform_initial_challenge(ctx, oracle, size, batch):
    """Config 1: new -> transform on the deterministic initial accumulator; response hash oracle == GPU."""
    from phase2_bn254_b200.powersoftau import BatchedAccumulator, CeremonyParams, PrivateKey, calculate_hash
    params = CeremonyParams(size, batch)
    challenge = oracle.pot_generate_initial(size)
    assert len(challenge) == params.accumulator_size
    response = np.zeros(params.contribution_size, dtype=np.uint8)
    response[:64] = np.frombuffer(calculate_hash(np.frombuffer(challenge, dtype=np.uint8)), dtype=np.uint8)
    BatchedAccumulator.transform(np.frombuffer(challenge, dtype=np.uint8), response, False, True, False,
                                 PrivateKey(TAU, ALPHA, BETA), params, ctx=ctx)
    exp = oracle.pot_transform(challenge, size, batch, be(TAU), be(ALPHA), be(BETA), threads=8)
    acc_len = params.contribution_size - params.public_key_size
    assert len(exp) == acc_len
    assert hashlib.blake2b(response[:acc_len].tobytes()).hexdigest() == hashlib.blake2b(exp).hexdigest()


def test_transform_chain_and_modes(ctx, oracle):
    """Second contribution on a non-trivial accumulator, every compression combination, checked input, 2 shards."""
    from phase2_bn254_b200.powersoftau import BatchedAccumulator, CeremonyParams, PrivateKey
    size, batch = 5, 8
    params = CeremonyParams(size, batch)
    ch0 = oracle.pot_generate_initial(size)
    ch1 = oracle.pot_transform(ch0, size, batch, be(TAU), be(ALPHA), be(BETA), out_compressed=False, threads=8)
    ch1c = oracle.pot_transform(ch0, size, batch, be(TAU), be(ALPHA), be(BETA), out_compressed=True, threads=8)
    key = PrivateKey(ALPHA, BETA, TAU)
    for in_c, src in ((False, ch1), (True, ch1c)):
        for out_c in (False, True):
            for check in (False, True):
                exp = oracle.pot_transform(src, size, batch, be(ALPHA), be(BETA), be(TAU), in_c, out_c, check, threads=8)
                out = np.zeros(len(exp), dtype=np.uint8)
                # two shards writing disjoint ranges of the same response (the multi-GPU decomposition)
                for shard in range(2):
                    BatchedAccumulator.transform(np.frombuffer(src, dtype=np.uint8), out, in_c, out_c, check, key,
                                                 params, ctx=ctx, shard_index=shard, shard_count=2)
                assert out[64:].tobytes() == exp[64:], (in_c, out_c, check)


def test_transform_rejects_infinity_and_garbage(ctx, oracle):
    from phase2_bn254_b200.powersoftau import (BatchedAccumulator, CeremonyParams, DeserializationError, PrivateKey)
    size = 3
    params = CeremonyParams(size, 4)
    ch = bytearray(oracle.pot_generate_initial(size))
    out = np.zeros(params.contribution_size, dtype=np.uint8)
    bad = bytearray(ch)
    bad[64 + 2 * 64: 64 + 3 * 64] = bytes([0x40]) + bytes(63)
    with pytest.raises(DeserializationError) as e:
        BatchedAccumulator.transform(np.frombuffer(bytes(bad), dtype=np.uint8), out, False, True, False,
                                     PrivateKey(TAU, ALPHA, BETA), params, ctx=ctx)
    assert e.value.kind == "PointAtInfinity"
    with pytest.raises(AssertionError):
        BatchedAccumulator.transform(np.frombuffer(bytes(ch), dtype=np.uint8), out, False, True, False,
                                     PrivateKey(TAU, 0, BETA), params, ctx=ctx)      # alpha = 0 -> infinity produced
    with pytest.raises(ValueError):
        BatchedAccumulator.transform(np.frombuffer(bytes(ch[:-1]), dtype=np.uint8), out, False, True, False,
                                     PrivateKey(TAU, ALPHA, BETA), params, ctx=ctx)


def synthetic_params(oracle, m, n_contrib=0, seed=5):
    """A synthetic MPCParameters file: h = m-1 points, l = m points, a/b_g1/b_g2 = 3, ic = 2 (config 3 shape)."""
    g1 = lambda n, s: random_points(oracle, 0, n, seed + s)
    g2 = lambda n, s: random_points(oracle, 1, n, seed + s)
    body = g1(1, 1) + g1(1, 2) + g2(1, 3) + g2(1, 4) + g1(1, 5) + g2(1, 6)
    for n, s, grp in ((2, 7, 0), (m - 1, 8, 0), (m, 9, 0), (3, 10, 0), (3, 11, 0), (3, 12, 1)):
        body += struct.pack(">I", n) + (g2(n, s) if grp else g1(n, s))
    body += hashlib.blake2b(body).digest()
    contribs = b"".join(g1(1, 20 + i) + g1(1, 30 + i) + g1(1, 40 + i) + g2(1, 50 + i) + hashlib.blake2b(bytes([i])).digest()
                        for i in range(n_contrib))
    return body + struct.pack(">I", n_contrib) + contribs


@pytest.mark.parametrize("m,n_contrib", [(8, 0), (300, 2)])
def test_phase2_contribute(ctx, oracle, m, n_contrib):
    from phase2_bn254_b200.phase2 import MPCParameters
    buf = synthetic_params(oracle, m, n_contrib)
    delta = 0x0fedcba987654321fedcba987654321fedcba987654321fedcba987654321 % R_MOD
    s = random_points(oracle, 0, 1, 77)
    r = random_points(oracle, 1, 1, 78)
    exp_file, exp_hash = oracle.phase2_contribute(buf, be(delta), s, r, threads=8)
    p = MPCParameters.read(buf)
    got_hash = p.contribute(delta, s, r_g2=r, ctx=ctx)
    assert got_hash == exp_hash
    assert p.data.tobytes() == exp_file
    # a second contribution on top (the contributions list grows)
    exp2, h2 = oracle.phase2_contribute(exp_file, be(delta + 1), s, r, threads=8)
    assert p.contribute(delta + 1, s, r_g2=r, ctx=ctx) == h2 and p.data.tobytes() == exp2


def test_transform_with_g2_subgroup_flag(ctx, oracle):
    """The opt-in endomorphism path for the TauG2 / BetaG2 sections gives the same response bytes."""
    from phase2_bn254_b200.powersoftau import BatchedAccumulator, CeremonyParams, PrivateKey
    size, batch = 6, 16
    params = CeremonyParams(size, batch)
    ch0 = oracle.pot_generate_initial(size)
    ch1 = oracle.pot_transform(ch0, size, batch, be(TAU), be(ALPHA), be(BETA), out_compressed=False, threads=8)
    exp = oracle.pot_transform(ch1, size, batch, be(BETA), be(TAU), be(ALPHA), threads=8)
    out = np.zeros(params.contribution_size, dtype=np.uint8)
    BatchedAccumulator.transform(np.frombuffer(ch1, dtype=np.uint8), out, False, True, False, PrivateKey(BETA, TAU, ALPHA),
                                 params, ctx=ctx, g2_in_subgroup=True)
    assert out[64:len(exp)].tobytes() == exp[64:]


@pytest.mark.parametrize("group", [0, 1])
def test_bulk_codec(ctx, oracle, group):
    """p2b_g{1,2}_recode: compress / decompress / checked deserialisation without a scalar multiplication."""
    from phase2_bn254_b200 import lib
    size = 128 if group else 64
    n = 500
    pts = bytearray(random_points(oracle, group, n, seed=81))
    pts[size * 7: size * 8] = bytes([0x40]) + bytes(size - 1)           # infinity is a valid encoding
    pts = bytes(pts)
    comp = oracle.batch_mul(group, pts, be(1), 0, 1, threads=8)
    assert ctx.recode(group, pts, 0, 1).tobytes() == comp
    assert ctx.recode(group, comp, 1, 0, lib.CHECK_INPUT).tobytes() == pts
    assert ctx.recode(group, comp, 1, 1).tobytes() == comp
    assert ctx.recode(group, pts, 0, 0, lib.CHECK_INPUT).tobytes() == pts
    with pytest.raises(lib.P2BError) as e:
        ctx.recode(group, pts, 0, 1, lib.REJECT_INFINITY)
    assert e.value.code == lib.EINFINITY_IN and e.value.index == 7
    bad = bytearray(pts); bad[size * 11 + size - 1] ^= 1
    with pytest.raises(lib.P2BError) as e:
        ctx.recode(group, bytes(bad), 0, 1, lib.CHECK_INPUT)
    assert e.value.code == lib.EDECODE and e.value.sub == lib.DEC_NOT_ON_CURVE and e.value.index == 11
    assert ctx.recode(group, bytes(bad), 0, 0).tobytes() == bytes(bad)   # unchecked accepts, like into_affine_unchecked
    # a compressed x with no square root
    x = bytearray(comp[: size // 2])
    for t in range(1, 50):
        x[-1] = (x[-1] + t) & 0xff
        try:
            oracle.point_recode(group, bytes(x), 1, 0)
        except oracle.OracleError:
            break
    with pytest.raises(lib.P2BError) as e:
        ctx.recode(group, bytes(x), 1, 0)
    assert e.value.code == lib.EDECODE and e.value.sub == lib.DEC_NOT_ON_CURVE


def test_pot_decompress(ctx, oracle):
    """BatchedAccumulator::decompress: compressed response -> next challenge, two shards, vs the oracle."""
    from phase2_bn254_b200.powersoftau import BatchedAccumulator, CeremonyParams, DeserializationError
    size, batch = 5, 8
    params = CeremonyParams(size, batch)
    ch0 = oracle.pot_generate_initial(size)
    resp = oracle.pot_transform(ch0, size, batch, be(TAU), be(ALPHA), be(BETA), threads=8)            # compressed
    nxt = oracle.pot_transform(ch0, size, batch, be(TAU), be(ALPHA), be(BETA), out_compressed=False, threads=8)
    out = np.zeros(params.accumulator_size, dtype=np.uint8)
    for shard in range(2):
        BatchedAccumulator.decompress(np.frombuffer(resp, dtype=np.uint8), out, True, params, ctx=ctx, shard_index=shard,
                                      shard_count=2)
    assert out[64:].tobytes() == nxt[64:]
    bad = bytearray(resp)
    bad[64 + 32 * 3] = 0x40                                             # infinity flag with stray bits
    with pytest.raises(DeserializationError):
        BatchedAccumulator.decompress(np.frombuffer(bytes(bad), dtype=np.uint8), out, False, params, ctx=ctx)


def test_transform_composition_property_large(ctx):
    """Size-independent check at 2^14 (82k G1 + 16k G2 points): contributing (tau1, alpha1, beta1) and then
    (tau2, alpha2, beta2) gives exactly the response of one contribution with the products of the secrets."""
    import hashlib
    from phase2_bn254_b200.powersoftau import BatchedAccumulator, CeremonyParams, PrivateKey
    size = 14
    params = CeremonyParams(size, 1024)
    ch0 = np.zeros(params.accumulator_size, dtype=np.uint8)
    o = 64
    g1a, g2a = np.frombuffer(G1_GEN, dtype=np.uint8), np.frombuffer(G2_GEN, dtype=np.uint8)
    for cnt, g in ((params.powers_g1_length, g1a), (params.powers_length, g2a), (params.powers_length, g1a),
                   (params.powers_length, g1a), (1, g2a)):
        ch0[o:o + cnt * g.size].reshape(cnt, g.size)[:] = g
        o += cnt * g.size
    k1, k2 = PrivateKey(TAU, ALPHA, BETA), PrivateKey(BETA + 5, TAU + 7, ALPHA + 11)
    k12 = PrivateKey(k1.tau * k2.tau % R_MOD, k1.alpha * k2.alpha % R_MOD, k1.beta * k2.beta % R_MOD)
    ch1 = np.zeros(params.accumulator_size, dtype=np.uint8)
    BatchedAccumulator.transform(ch0, ch1, False, False, False, k1, params, ctx=ctx)
    r2 = np.zeros(params.contribution_size, dtype=np.uint8)
    BatchedAccumulator.transform(ch1, r2, False, True, True, k2, params, ctx=ctx)
    r12 = np.zeros(params.contribution_size, dtype=np.uint8)
    BatchedAccumulator.transform(ch0, r12, False, True, False, k12, params, ctx=ctx)
    end = params.contribution_size - params.public_key_size
    assert hashlib.blake2b(r2[64:end].tobytes()).digest() == hashlib.blake2b(r12[64:end].tobytes()).digest()
    # and the endomorphism path for G2 agrees on the whole file
    r12s = np.zeros(params.contribution_size, dtype=np.uint8)
    BatchedAccumulator.transform(ch0, r12s, False, True, False, k12, params, ctx=ctx, g2_in_subgroup=True)
    assert np.array_equal(r12s[64:end], r12[64:end])


def test_transform_2_14_matches_oracle(ctx, oracle):
    """Whole-file parity at 2^14 (the largest size the CPU oracle finishes in seconds)."""
    from phase2_bn254_b200.powersoftau import BatchedAccumulator, CeremonyParams, PrivateKey
    size, batch = 14, 512
    params = CeremonyParams(size, batch)
    ch0 = oracle.pot_generate_initial(size)
    exp = oracle.pot_transform(ch0, size, batch, be(TAU), be(ALPHA), be(BETA), threads=16)
    out = np.zeros(params.contribution_size, dtype=np.uint8)
    BatchedAccumulator.transform(np.frombuffer(ch0, dtype=np.uint8), out, False, True, False, PrivateKey(TAU, ALPHA, BETA),
                                 params, ctx=ctx)
    assert out[64:len(exp)].tobytes() == exp[64:]


def test_mpc_parameters_read_checked(ctx):
    """MPCParameters::read: checked deserialisation of every section on the GPU (groth16/mod.rs:287-383)."""
    import json
    import os
    from phase2_bn254_b200.phase2 import MPCParameters, params_layout
    gold = json.load(open(os.path.join(os.path.dirname(__file__), "golden", "vectors.json")))["phase2"]
    good = bytes.fromhex(gold["c1"]["out"])
    mp = MPCParameters.read(good, False, True, ctx=ctx)                       # h[2] is infinity: allowed by default
    assert mp.data.tobytes() == good
    with pytest.raises(IOError) as e:
        MPCParameters.read(good, True, True, ctx=ctx)
    assert "point at infinity in h[2]" in str(e.value)
    lay = params_layout(good)
    bad = bytearray(good)
    bad[lay["b_g2"][0] + 128 + 127] ^= 1                                       # b_g2[1] off the curve
    with pytest.raises(IOError) as e:
        MPCParameters.read(bytes(bad), False, True, ctx=ctx)
    assert "NotOnCurve in b_g2[1]" in str(e.value)
    MPCParameters.read(bytes(bad), False, False, ctx=ctx)                      # unchecked accepts it (into_affine_unchecked)
    bad = bytearray(good)
    bad[lay["a"][0]: lay["a"][0] + 32] = (2**256 - 1).to_bytes(32, "big")      # flag bits + coordinate >= q
    with pytest.raises(IOError):
        MPCParameters.read(bytes(bad), False, False, ctx=ctx)


def test_phase2_contribute_sharded_equals_whole(ctx, oracle):
    """H and L split into contiguous ranges (one per GPU): the shards, run one after the other into the same output buffer,
    produce the bytes of the unsharded call, and every shard returns the same contribution hash."""
    buf = synthetic_params(oracle, 301, 1)
    delta = be(0x0fedcba987654321fedcba987654321fedcba987654321fedcba987654321 % R_MOD)
    s = random_points(oracle, 0, 1, 77)
    r = random_points(oracle, 1, 1, 78)
    exp_file, exp_hash = oracle.phase2_contribute(buf, delta, s, r, threads=8)
    b = np.frombuffer(buf, dtype=np.uint8)
    for shards in (2, 3, 8):
        out = np.zeros(b.size + 384, dtype=np.uint8)
        hashes = {ctx.phase2_contribute(b, delta, s, r, out=out, shard_index=i, shard_count=shards)[1] for i in reversed(range(shards))}
        assert hashes == {exp_hash}
        assert out.tobytes() == exp_file
