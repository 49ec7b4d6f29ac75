"""GPU parity for the one-pass verifier combinations: p2b_g{1,2}_msm_pair (merge_pairs, phase2/src/utils.rs:59-105,
powersoftau/src/utils.rs:112-130) and p2b_g{1,2}_power_pairs (powersoftau/src/utils.rs:133-135) against two oracle MSMs
over the same coefficients; the device-generated coefficients against the python ChaCha20 (pinned to RFC 7539 in
tests/test_chacha_ref.py)."""
import numpy as np
import pytest

from util import G1_GEN, G2_GEN, R_MOD, be, device_scalars, random_points, random_scalars

pytestmark = pytest.mark.gpu
SEED = bytes((7 * i + 3) & 0xff for i in range(32))


@pytest.mark.parametrize("bits", [253, 248, 128, 64])
def test_random_scalars_match_chacha20(ctx, bits):
    for n, first in ((1, 0), (2, 0), (5, 0), (1000, 0), (7, 3), (6, 1 << 33)):
        assert ctx.random_scalars(SEED, n, bits, first).tobytes() == device_scalars(SEED, n, bits, first), (n, first)


@pytest.mark.parametrize("group", [0, 1])
@pytest.mark.parametrize("n", [1, 2, 33, 700])
def test_msm_pair_given_scalars(ctx, oracle, group, n):
    a, b = random_points(oracle, group, n, seed=1100 + n), random_points(oracle, group, n, seed=1200 + n)
    sc = random_scalars(n, seed=1300 + n)
    ra, rb = ctx.msm_pair(group, a, b, sc)
    assert ra == oracle.msm(group, a, sc, threads=8) and rb == oracle.msm(group, b, sc, threads=8)


@pytest.mark.parametrize("group", [0, 1])
@pytest.mark.parametrize("bits", [253, 128])
def test_msm_pair_device_coefficients(ctx, oracle, group, bits):
    n = 900
    a, b = random_points(oracle, group, n, seed=1400), random_points(oracle, group, n, seed=1401)
    ra, rb = ctx.msm_pair(group, a, b, None, SEED, bits)
    rho = device_scalars(SEED, n, bits)
    assert ra == oracle.msm(group, a, rho, threads=8) and rb == oracle.msm(group, b, rho, threads=8)


@pytest.mark.parametrize("group", [0, 1])
def test_power_pairs_matches_two_msms_and_the_ratio(ctx, oracle, group):
    size = 128 if group else 64
    n = 513
    tau = be(0x1234567 ** 9 % R_MOD)
    v = oracle.batch_mul_powers(group, (G2_GEN if group else G1_GEN) * n, tau, None, 0, threads=8)
    for bits in (253, 128):
        a, b = ctx.power_pairs(group, v, None, SEED, bits)
        rho = device_scalars(SEED, n - 1, bits)
        assert a == oracle.msm(group, v[: (n - 1) * size], rho, threads=8)
        assert b == oracle.msm(group, v[size:], rho, threads=8)
        assert b == oracle.point_mul(group, a, tau)                       # consecutive elements have ratio tau
    sc = random_scalars(n - 1, seed=1500)
    a, b = ctx.power_pairs(group, v, sc)
    assert a == oracle.msm(group, v[: (n - 1) * size], sc, threads=8) and b == oracle.msm(group, v[size:], sc, threads=8)
    # compressed input (the response file's encoding): decompressed on the device
    comp = oracle.batch_mul(group, v, be(1), 0, 1, threads=8)
    from phase2_bn254_b200 import lib
    assert ctx.power_pairs(group, comp, sc, in_enc=lib.ENC_COMPRESSED, flags=lib.REJECT_INFINITY) == (a, b)
    # degenerate sizes: one point -> no terms -> (infinity, infinity); two points -> one term
    inf = bytes([0x40]) + bytes(size - 1)
    assert ctx.power_pairs(group, v[:size], None, SEED, 253) == (inf, inf)
    a2, b2 = ctx.power_pairs(group, v[: 2 * size], be(5))
    assert a2 == oracle.point_mul(group, v[:size], be(5)) and b2 == oracle.point_mul(group, v[size: 2 * size], be(5))


def test_pair_streamed_chunks_and_errors(ctx, oracle, monkeypatch):
    from phase2_bn254_b200 import lib
    n = 3000
    a, b = random_points(oracle, 0, n, seed=1600), random_points(oracle, 0, n, seed=1601)
    rho = device_scalars(SEED, n, 253)
    exp = (oracle.msm(0, a, rho, threads=8), oracle.msm(0, b, rho, threads=8))
    v = a + b[:64]
    exp_pp = (oracle.msm(0, v[: n * 64], rho, threads=8), oracle.msm(0, v[64:], rho, threads=8))
    for chunk in (700, 1024, 2999):
        monkeypatch.setenv("P2B_MSM_STREAM_CHUNK", str(chunk))          # several chunks continue the same two bucket sets;
        assert ctx.msm_pair(0, a, b, None, SEED, 253) == exp              # the generated coefficients continue across chunks
        assert ctx.power_pairs(0, v, None, SEED, 253) == exp_pp           # chunks overlap by one point
    monkeypatch.delenv("P2B_MSM_STREAM_CHUNK")
    # decode errors: index of the failing point; curve check and infinity rejection on request
    bad = bytearray(b)
    bad[64 * 1234] |= 0x80
    with pytest.raises(lib.P2BError) as e:
        ctx.msm_pair(0, a, bytes(bad), None, SEED, 253)
    assert e.value.code == lib.EDECODE and e.value.index == 1234
    off = bytearray(a)
    off[64 * 77 + 63] ^= 1
    ctx.msm_pair(0, bytes(off), b, None, SEED, 253)                       # unchecked: summed as given
    with pytest.raises(lib.P2BError) as e:
        ctx.msm_pair(0, bytes(off), b, None, SEED, 253, flags=lib.CHECK_INPUT)
    assert e.value.code == lib.EDECODE and e.value.sub == lib.DEC_NOT_ON_CURVE and e.value.index == 77
    inf = bytearray(a)
    inf[64 * 5: 64 * 6] = bytes([0x40]) + bytes(63)
    ra, _ = ctx.msm_pair(0, bytes(inf), b, None, SEED, 253)               # infinity contributes nothing
    assert ra == oracle.msm(0, bytes(inf), rho, threads=8)
    with pytest.raises(lib.P2BError) as e:
        ctx.msm_pair(0, bytes(inf), b, None, SEED, 253, flags=lib.REJECT_INFINITY)
    assert e.value.code == lib.EINFINITY_IN and e.value.index == 5
    # a coefficient above the promised bound is refused, not silently truncated
    with pytest.raises(lib.P2BError) as e:
        ctx.msm_pair(0, a[:128], b[:128], be(1) + be(1 << 130), scalar_bits=128)
    assert e.value.code == lib.EARG and e.value.index == 1
    with pytest.raises(lib.P2BError):
        ctx.msm_pair(0, a, b, None, None, 253)                            # neither scalars nor a seed


def test_pair_2p20_matches_oracle(ctx, oracle):
    """power_pairs at a verifier chunk size (2^20 points, 128-bit coefficients) against the oracle."""
    n = 1 << 20
    v = ctx.batch_mul_powers(0, np.tile(np.frombuffer(G1_GEN, dtype=np.uint8), n), be(0x1234567 ** 9 % R_MOD), None, 0)
    a, b = ctx.power_pairs(0, v, None, SEED, 128)
    rho = np.frombuffer(ctx.random_scalars(SEED, n - 1, 128).tobytes(), dtype=np.uint8)
    assert rho[: 32 * 64].tobytes() == device_scalars(SEED, 64, 128)
    assert a == oracle.msm(0, v[: (n - 1) * 64], rho, threads=32)
    assert b == oracle.msm(0, v[64:], rho, threads=32)
