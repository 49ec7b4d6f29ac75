"""GPU parity: batched scalar multiplication through the C ABI vs the CPU oracle (bit-exact)."""
import numpy as np
import pytest

from util import EDGE_SCALARS, G1_GEN, G2_GEN, R_MOD, be, random_points, random_scalars

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("group", [0, 1])
@pytest.mark.parametrize("n", [1, 31, 255, 1024])
def test_batch_mul_per_point_scalars(ctx, oracle, group, n):
    pts = random_points(oracle, group, n, seed=100 + n)
    sc = random_scalars(n, seed=200 + n)
    for out_enc in (0, 1):
        got = ctx.batch_mul(group, pts, sc, 0, out_enc).tobytes()
        exp = oracle.batch_mul(group, pts, sc, 0, out_enc, threads=8)
        assert got == exp


@pytest.mark.parametrize("group", [0, 1])
def test_batch_mul_edge_scalars(ctx, oracle, group):
    n = len(EDGE_SCALARS)
    pts = random_points(oracle, group, n, seed=7)
    sc = b"".join(be(k) for k in EDGE_SCALARS)
    got = ctx.batch_mul(group, pts, sc).tobytes()
    exp = oracle.batch_mul(group, pts, sc, threads=4)
    assert got == exp


@pytest.mark.parametrize("group", [0, 1])
def test_batch_mul_broadcast_and_infinity(ctx, oracle, group):
    n = 300
    pts = bytearray(random_points(oracle, group, n, seed=9))
    size = 128 if group else 64
    pts[5 * size:6 * size] = bytes([0x40]) + bytes(size - 1)          # a point at infinity is tolerated (phase-2 shape)
    k = be(0x1234567890abcdef1234567890abcdef1234567890abcdef % R_MOD)
    got = ctx.batch_mul(group, bytes(pts), k).tobytes()
    exp = oracle.batch_mul(group, bytes(pts), k, threads=8)
    assert got == exp
    assert got[5 * size] == 0x40


@pytest.mark.parametrize("group", [0, 1])
def test_batch_mul_compressed_input(ctx, oracle, group):
    n = 200
    pts = random_points(oracle, group, n, seed=11)
    sc = random_scalars(n, seed=12)
    comp = oracle.batch_mul(group, pts, be(1), 0, 1, threads=8)       # compress
    got = ctx.batch_mul(group, comp, sc, 1, 1).tobytes()
    exp = oracle.batch_mul(group, comp, sc, 1, 1, threads=8)
    assert got == exp


@pytest.mark.parametrize("group", [0, 1])
def test_batch_mul_powers(ctx, oracle, group):
    n = 700
    pts = random_points(oracle, group, n, seed=13)
    tau, coeff = be(0x1111111111111111111111111111111111111111111111111111111111111111 % R_MOD), be(R_MOD - 5)
    for start, cf in ((0, None), (12345, coeff), ((1 << 27) + 3, coeff)):
        got = ctx.batch_mul_powers(group, pts, tau, cf, start, 0, 1).tobytes()
        exp = oracle.batch_mul_powers(group, pts, tau, cf, start, 0, 1, threads=8)
        assert got == exp


def test_errors(ctx, oracle):
    from phase2_bn254_b200 import lib
    n = 64
    pts = bytearray(random_points(oracle, 0, n, seed=21))
    sc = random_scalars(n, seed=22)
    bad = bytearray(pts)
    bad[10 * 64 + 63] ^= 1                                           # off the curve
    with pytest.raises(lib.P2BError) as e:
        ctx.batch_mul(0, bytes(bad), sc, flags=lib.CHECK_INPUT)
    assert e.value.code == lib.EDECODE and e.value.sub == lib.DEC_NOT_ON_CURVE and e.value.index == 10
    # unchecked: the reference computes on whatever it decoded; so do we (slow generic path), bit-exact
    got = ctx.batch_mul(0, bytes(bad), sc).tobytes()
    assert got == oracle.batch_mul(0, bytes(bad), sc, threads=4)
    bad = bytearray(pts)
    bad[3 * 64] |= 0x80                                              # sign bit on an uncompressed point
    with pytest.raises(lib.P2BError) as e:
        ctx.batch_mul(0, bytes(bad), sc)
    assert e.value.code == lib.EDECODE and e.value.sub == lib.DEC_UNEXPECTED_INFORMATION and e.value.index == 3
    bad = bytearray(pts)
    bad[7 * 64:7 * 64 + 32] = b"\x3f" + b"\xff" * 31                 # x >= q
    with pytest.raises(lib.P2BError) as e:
        ctx.batch_mul(0, bytes(bad), sc)
    assert e.value.code == lib.EDECODE and e.value.sub == lib.DEC_COORDINATE and e.value.index == 7
    bad = bytearray(pts)
    bad[2 * 64:3 * 64] = bytes([0x40]) + bytes(63)
    with pytest.raises(lib.P2BError) as e:
        ctx.batch_mul(0, bytes(bad), sc, flags=lib.REJECT_INFINITY)
    assert e.value.code == lib.EINFINITY_IN and e.value.index == 2
    zero = bytearray(sc)
    zero[32 * 9:32 * 10] = bytes(32)                                 # scalar 0 -> produced infinity
    with pytest.raises(lib.P2BError) as e:
        ctx.batch_mul(0, bytes(pts), bytes(zero), flags=lib.REJECT_INFINITY)
    assert e.value.code == lib.EINFINITY_OUT and e.value.index == 9
    g2pts = bytearray(random_points(oracle, 1, 4, seed=23))
    g2pts[128] |= 0x80
    with pytest.raises(lib.P2BError) as e:
        ctx.batch_mul(1, bytes(g2pts), random_scalars(4, 1))
    assert e.value.code == lib.EDECODE and e.value.sub == lib.DEC_UNEXPECTED_COMPRESSION_MODE and e.value.index == 1


def test_empty(ctx):
    assert ctx.batch_mul(0, b"", be(5)).size == 0


def test_linearity_large(ctx, oracle):
    """Size-independent property at a size the oracle would take minutes for: [a]P + ... checked via
    [a]([b]P) == [ab]P on 2^17 points, plus an oracle spot check of 64 of them."""
    n = 1 << 17
    base = np.frombuffer(random_points(oracle, 0, 256, seed=31), dtype=np.uint8)
    pts = np.tile(base, n // 256)
    a, b = 0x0123456789abcdef0123456789abcdef0123456789abcdef0123456789abcdef % R_MOD, R_MOD - 12345
    p1 = ctx.batch_mul(0, pts, be(a))
    p2 = ctx.batch_mul(0, p1, be(b))
    p3 = ctx.batch_mul(0, pts, be(a * b % R_MOD))
    assert np.array_equal(p2, p3)
    assert p3[:64 * 64].tobytes() == oracle.batch_mul(0, pts[:64 * 64].tobytes(), be(a * b % R_MOD), threads=8)


def test_g2_subgroup_flag_uses_endomorphism_and_matches(ctx, oracle):
    """P2B_G2_SUBGROUP: the opt-in endomorphism split for G2 inputs known to lie in the order-r subgroup."""
    from phase2_bn254_b200 import lib
    n = len(EDGE_SCALARS) + 300
    pts = random_points(oracle, 1, n, seed=31)
    sc = b"".join(be(k) for k in EDGE_SCALARS) + random_scalars(300, seed=32)
    exp = oracle.batch_mul(1, pts, sc, 0, 1, threads=8)
    assert ctx.batch_mul(1, pts, sc, 0, 1, flags=lib.G2_SUBGROUP).tobytes() == exp
    assert ctx.batch_mul(1, pts, sc, 0, 1).tobytes() == exp
    tau = be(0x1234567 ** 9 % R_MOD)
    exp = oracle.batch_mul_powers(1, pts, tau, None, 5, threads=8)
    assert ctx.batch_mul_powers(1, pts, tau, None, 5, flags=lib.G2_SUBGROUP | lib.REJECT_INFINITY).tobytes() == exp


@pytest.mark.parametrize("group", [0, 1])
def test_empty_and_ragged_sizes(ctx, oracle, group):
    """n = 0 is a no-op everywhere; sizes around the block / warp / chunk boundaries."""
    assert ctx.batch_mul(group, b"", be(5)).size == 0
    assert ctx.batch_mul_powers(group, b"", be(7)).size == 0
    assert ctx.recode(group, b"", 0, 1).size == 0
    for n in (33, 127, 129, 257, 513):
        pts = random_points(oracle, group, n, seed=500 + n)
        sc = random_scalars(n, seed=600 + n)
        assert ctx.batch_mul(group, pts, sc, 0, 1).tobytes() == oracle.batch_mul(group, pts, sc, 0, 1, threads=8)


def test_non_canonical_scalar_is_rejected(ctx, oracle):
    from phase2_bn254_b200 import lib
    pts = random_points(oracle, 0, 4, seed=1)
    with pytest.raises(lib.P2BError) as e:
        ctx.batch_mul(0, pts, be(1) * 2 + be(R_MOD) + be(1))
    assert e.value.code == lib.EARG and e.value.index == 2
    with pytest.raises(lib.P2BError):
        ctx.batch_mul(0, pts, be(R_MOD))
    with pytest.raises(lib.P2BError):
        ctx.batch_mul(0, pts, be(1) * 3)              # n_scalars must be 1 or n
