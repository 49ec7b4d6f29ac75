"""GPU: the verifier mirrors -- BatchedAccumulator.verify_transformation (batched_accumulator.rs:279-540) and phase2
verify_contribution (parameters.rs:722-855) -- on contributions produced by the GPU contribution path.  The random linear
combinations (power_pairs / merge_pairs) are Pippenger MSMs on the GPU, the same_ratio pairings run on the host.  This is
the reference's own end-to-end check (powersoftau/test.sh, phase2/test.sh: contribute, then the verifier must accept),
plus the tamper cases a verifier exists for."""
import hashlib
import struct

import numpy as np
import pytest

from util import G1_GEN, G2_GEN, R_MOD, be, random_points

pytestmark = pytest.mark.gpu


def _contribute_phase1(ctx, lib, challenge, params, seed, out_compressed=True, in_compressed=False):
    from phase2_bn254_b200.powersoftau import BatchedAccumulator, calculate_hash, keypair
    digest = calculate_hash(challenge)
    pub, priv = keypair(lib.ChaChaRng.from_digest(hashlib.sha256(seed).digest()), digest)
    n = params.contribution_size if out_compressed else params.accumulator_size + params.public_key_size
    response = np.zeros(n, dtype=np.uint8)
    response[:64] = np.frombuffer(digest, dtype=np.uint8)
    BatchedAccumulator.transform(challenge, response, in_compressed, out_compressed, False, priv, params, ctx=ctx)
    pub.write(response, out_compressed, params)
    return response, pub, digest


@pytest.mark.parametrize("size,batch", [(4, 4), (6, 16), (8, 256)])
def test_verify_transformation_chain(ctx, size, batch):
    """new -> contribute -> verify -> decompress -> contribute -> verify, as powersoftau/test.sh does."""
    from phase2_bn254_b200 import lib
    from phase2_bn254_b200.powersoftau import (BatchedAccumulator, CeremonyParams, PublicKey, calculate_hash,
                                               verify_transformation)
    params = CeremonyParams(size, batch)
    rng = np.random.default_rng(11)
    ch0 = np.zeros(params.accumulator_size, dtype=np.uint8)
    ch0[:64] = np.frombuffer(hashlib.blake2b(b"").digest(), dtype=np.uint8)
    BatchedAccumulator.generate_initial(ch0, False, params)
    rs1, pub1, d1 = _contribute_phase1(ctx, lib, ch0, params, b"first")
    assert PublicKey.read(rs1, True, params) == pub1
    assert verify_transformation(ch0, rs1, pub1, d1, False, True, False, True, params, ctx=ctx, rng=rng)
    assert verify_transformation(ch0, rs1, pub1, d1, False, True, True, True, params, ctx=ctx, rng=rng)
    # next challenge = decompressed response with the hash of the response in front (verify_transform_constrained.rs:207-229)
    ch1 = np.zeros(params.accumulator_size, dtype=np.uint8)
    ch1[:64] = np.frombuffer(calculate_hash(rs1), dtype=np.uint8)
    BatchedAccumulator.decompress(rs1, ch1, True, params, ctx=ctx)
    rs2, pub2, d2 = _contribute_phase1(ctx, lib, ch1, params, b"second")
    assert verify_transformation(ch1, rs2, pub2, d2, False, True, False, True, params, ctx=ctx, rng=rng)
    # compressed -> uncompressed variant of the same step verifies too (input read compressed from the response)
    rs2u, pub2u, d2u = _contribute_phase1(ctx, lib, rs1, params, b"second", out_compressed=False, in_compressed=True)
    assert verify_transformation(rs1, rs2u, pub2u, d2u, True, False, True, True, params, ctx=ctx, rng=rng)
    # -- rejections
    assert not verify_transformation(ch1, rs2, pub1, d2, False, True, False, True, params, ctx=ctx, rng=rng)   # wrong key
    assert not verify_transformation(ch1, rs2, pub2, d1, False, True, False, True, params, ctx=ctx, rng=rng)   # wrong digest
    assert not verify_transformation(ch0, rs2, pub2, d2, False, True, False, True, params, ctx=ctx, rng=rng)   # wrong predecessor
    sec = {"tau_g1": 64, "tau_g2": 64 + params.powers_g1_length * 32,
           "alpha_g1": 64 + params.powers_g1_length * 32 + params.powers_length * 64}
    sec["beta_g1"] = sec["alpha_g1"] + params.powers_length * 32
    other = ctx.recode(0, rs1[sec["tau_g1"] + 32 * 2: sec["tau_g1"] + 32 * 3], 1, 1).tobytes()     # a valid foreign G1 point
    for name, idx in (("tau_g1", 3), ("tau_g1", params.powers_g1_length - 1), ("tau_g1", params.powers_length),
                      ("alpha_g1", params.powers_length - 1), ("beta_g1", 1)):
        bad = rs2.copy()
        off = sec[name] + 32 * idx
        bad[off: off + 32] = np.frombuffer(other, dtype=np.uint8)
        assert not verify_transformation(ch1, bad, pub2, d2, False, True, False, True, params, ctx=ctx, rng=rng), (name, idx)
    bad = rs2.copy()
    off = sec["tau_g2"] + 64 * (params.powers_length - 1)
    bad[off: off + 64] = rs1[off: off + 64]                                                        # a valid foreign G2 point
    assert not verify_transformation(ch1, bad, pub2, d2, False, True, False, True, params, ctx=ctx, rng=rng)


def _initial_phase2_params(oracle, m, seed=31):
    """Serialized MPCParameters as MPCParameters::new leaves them: delta_g1 / delta_g2 are the generators, no contributions."""
    g1 = lambda n, s: random_points(oracle, 0, n, seed + s)
    g2 = lambda n, s: random_points(oracle, 1, n, seed + s)
    body = g1(1, 1) + g1(1, 2) + g2(1, 3) + g2(1, 4) + G1_GEN + G2_GEN
    for n, s, grp in ((2, 7, 0), (m - 1, 8, 0), (m, 9, 0), (3, 10, 0), (3, 11, 0), (3, 12, 1)):
        body += struct.pack(">I", n) + (g2(n, s) if grp else g1(n, s))
    body += hashlib.blake2b(body).digest()
    return body + struct.pack(">I", 0)


@pytest.mark.parametrize("m", [8, 700])
def test_verify_contribution_chain(ctx, oracle, m):
    from phase2_bn254_b200 import lib
    from phase2_bn254_b200.phase2 import (MPCParameters, VerificationError, keypair, params_layout, verify_contribution)
    rng = np.random.default_rng(12)
    p0 = MPCParameters.read(_initial_phase2_params(oracle, m), ctx=ctx)
    p1 = MPCParameters(p0.data.copy())
    # keypair() consumes the RNG exactly as contribute() does: the public key can be predicted from the same seed
    pk_expected, delta = keypair(lib.ChaChaRng([1, 2, 3, 4, 5, 6, 7, 8]), p0)
    h1 = p1.contribute(rng=lib.ChaChaRng([1, 2, 3, 4, 5, 6, 7, 8]), ctx=ctx)
    lay = params_layout(p1.data)
    assert lay["contributions"][1] == 1
    assert p1.data[lay["contributions"][0]:].tobytes() == pk_expected
    assert h1 == hashlib.blake2b(pk_expected).digest()
    assert verify_contribution(p0, p1, ctx=ctx, rng=rng) == h1
    p2 = MPCParameters(p1.data.copy())
    h2 = p2.contribute(rng=lib.ChaChaRng([9, 9, 9, 9, 9, 9, 9, 9]), ctx=ctx)
    assert verify_contribution(p1, p2, ctx=ctx, rng=rng) == h2
    # explicit secrets instead of the RNG (the older entry point of the mirror)
    p3 = MPCParameters(p2.data.copy())
    h3 = p3.contribute(12345678901234567890 % R_MOD, random_points(oracle, 0, 1, 99), hash_to_g2=lib.hash_to_g2, ctx=ctx)
    assert verify_contribution(p2, p3, ctx=ctx, rng=rng) == h3
    # -- rejections
    with pytest.raises(VerificationError):
        verify_contribution(p0, p2, ctx=ctx, rng=rng)                   # two contributions at once
    with pytest.raises(VerificationError):
        verify_contribution(p1, p1, ctx=ctx, rng=rng)
    foreign = random_points(oracle, 0, 1, 1234)
    for name, idx in (("h", 3), ("l", m - 1), ("a", 1), ("ic", 0)):
        bad = MPCParameters(p2.data.copy())
        off = params_layout(bad.data)[name][0] + 64 * idx
        bad.data[off: off + 64] = np.frombuffer(foreign, dtype=np.uint8)
        with pytest.raises(VerificationError):
            verify_contribution(p1, bad, ctx=ctx, rng=rng)
    bad = MPCParameters(p2.data.copy())                                  # delta_g2 not matching delta_g1
    off = params_layout(bad.data)["delta_g2"][0]
    bad.data[off: off + 128] = p1.data[off: off + 128]
    with pytest.raises(VerificationError):
        verify_contribution(p1, bad, ctx=ctx, rng=rng)
    bad = MPCParameters(p2.data.copy())                                  # transcript byte flipped
    bad.data[-1] ^= 1
    with pytest.raises(VerificationError):
        verify_contribution(p1, bad, ctx=ctx, rng=rng)
