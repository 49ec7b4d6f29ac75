"""GPU parity for P2B_ENC_RAW_MONT_LE through the C ABI -- the encoding the Rust shim of MPCParameters::contribute uses
(INTEGRATION.md section 4): RawEncodable::{into_raw_uncompressed_le, from_raw_uncompressed_le[_unchecked]}
(pairing/src/bn256/ec.rs:653-706, trait pairing/src/lib.rs:236-246): x || y as Montgomery limbs, little-endian, all-zero
bytes = the point at infinity.  Checked two ways: against the oracle's own raw codec, and against the explicit *R / /R
conversion of the wire (big-endian canonical) results in python integers."""
import numpy as np
import pytest

from util import EDGE_SCALARS, Q_MOD, R_MOD, be, random_points, random_scalars

pytestmark = pytest.mark.gpu
RAW = 2
R256 = 1 << 256


def wire_to_raw(pts):
    """uncompressed wire G1 points -> RawEncodable bytes, by integer arithmetic (x * R mod q, little-endian)."""
    out = bytearray()
    for i in range(0, len(pts), 64):
        p = pts[i: i + 64]
        if p[0] & 0x40:
            out += bytes(64)
            continue
        x, y = int.from_bytes(p[:32], "big"), int.from_bytes(p[32:], "big")
        out += (x * R256 % Q_MOD).to_bytes(32, "little") + (y * R256 % Q_MOD).to_bytes(32, "little")
    return bytes(out)


def raw_to_wire(raw):
    rinv = pow(R256, -1, Q_MOD)
    out = bytearray()
    for i in range(0, len(raw), 64):
        p = raw[i: i + 64]
        if not any(p):
            out += bytes([0x40]) + bytes(63)
            continue
        x, y = int.from_bytes(p[:32], "little"), int.from_bytes(p[32:], "little")
        out += (x * rinv % Q_MOD).to_bytes(32, "big") + (y * rinv % Q_MOD).to_bytes(32, "big")
    return bytes(out)


def test_raw_codec_matches_reference_layout(ctx, oracle):
    n = 300
    pts = bytearray(random_points(oracle, 0, n, seed=901))
    pts[64 * 5: 64 * 6] = bytes([0x40]) + bytes(63)
    pts = bytes(pts)
    raw = ctx.recode(0, pts, 0, RAW).tobytes()
    assert raw == wire_to_raw(pts)                                     # into_raw_uncompressed_le
    assert raw == oracle.batch_mul(0, pts, be(1), 0, RAW, threads=4)
    assert raw[64 * 5: 64 * 6] == bytes(64)                            # infinity = all-zero bytes
    from phase2_bn254_b200 import lib
    assert ctx.recode(0, raw, RAW, 0, lib.CHECK_INPUT).tobytes() == pts   # from_raw_uncompressed_le (checked)
    assert ctx.recode(0, raw, RAW, 1).tobytes() == oracle.batch_mul(0, pts, be(1), 0, 1, threads=4)


@pytest.mark.parametrize("n", [1, 33, 1000])
def test_raw_batch_mul_per_point_and_broadcast(ctx, oracle, n):
    pts = random_points(oracle, 0, n, seed=910 + n)
    raw = wire_to_raw(pts)
    sc = random_scalars(n, seed=920 + n)
    exp = oracle.batch_mul(0, pts, sc, threads=8)
    got = ctx.batch_mul(0, raw, sc, RAW, RAW).tobytes()
    assert got == oracle.batch_mul(0, raw, sc, RAW, RAW, threads=8)
    assert raw_to_wire(got) == exp
    k = be(0x1234567890abcdef1234567890abcdef1234567890abcdef % R_MOD)    # the phase-2 shape: one scalar for all points
    got = ctx.batch_mul(0, raw, k, RAW, RAW).tobytes()
    assert raw_to_wire(got) == oracle.batch_mul(0, pts, k, threads=8)
    # mixed encodings in / out
    assert ctx.batch_mul(0, raw, k, RAW, 0).tobytes() == oracle.batch_mul(0, pts, k, threads=8)
    assert ctx.batch_mul(0, pts, k, 0, RAW).tobytes() == wire_to_raw(oracle.batch_mul(0, pts, k, threads=8))


def test_raw_edge_scalars_and_infinity(ctx, oracle):
    n = len(EDGE_SCALARS)
    pts = bytearray(random_points(oracle, 0, n, seed=931))
    pts[64 * 2: 64 * 3] = bytes([0x40]) + bytes(63)
    pts = bytes(pts)
    raw = wire_to_raw(pts)
    sc = b"".join(be(k) for k in EDGE_SCALARS)
    got = ctx.batch_mul(0, raw, sc, RAW, RAW).tobytes()
    assert raw_to_wire(got) == oracle.batch_mul(0, pts, sc, threads=4)
    assert got[:64] == bytes(64) and got[128:192] == bytes(64)            # k = 0 and the infinity input -> all-zero
    # all-zero input buffer = n points at infinity (flags 0: tolerated, phase2/src/parameters.rs:467-469)
    assert ctx.batch_mul(0, bytes(64 * 7), be(5), RAW, RAW).tobytes() == bytes(64 * 7)
    from phase2_bn254_b200 import lib
    with pytest.raises(lib.P2BError) as e:                                # phase-1 semantics reject it
        ctx.batch_mul(0, raw, be(5), RAW, RAW, flags=lib.REJECT_INFINITY)
    assert e.value.code == lib.EINFINITY_IN and e.value.index == 2


def test_raw_decode_errors(ctx, oracle):
    """Fq::from_raw_repr rejects limbs >= q (CoordinateDecodingError); the checked variant rejects off-curve points."""
    from phase2_bn254_b200 import lib
    raw = bytearray(wire_to_raw(random_points(oracle, 0, 8, seed=941)))
    bad = bytearray(raw)
    bad[64 * 3: 64 * 3 + 32] = Q_MOD.to_bytes(32, "little")               # x = q: not in the field
    with pytest.raises(lib.P2BError) as e:
        ctx.batch_mul(0, bytes(bad), be(3), RAW, RAW)
    assert e.value.code == lib.EDECODE and e.value.sub == lib.DEC_COORDINATE and e.value.index == 3
    bad = bytearray(raw)
    bad[64 * 6 + 32] ^= 1                                                 # y perturbed: off the curve
    with pytest.raises(lib.P2BError) as e:
        ctx.batch_mul(0, bytes(bad), be(3), RAW, RAW, flags=lib.CHECK_INPUT)
    assert e.value.code == lib.EDECODE and e.value.sub == lib.DEC_NOT_ON_CURVE and e.value.index == 6
    with pytest.raises(oracle.OracleError) as oe:
        oracle.batch_mul(0, bytes(bad), be(3), RAW, RAW, checked=True)
    assert oe.value.code == oracle.EDECODE and oe.value.sub == oracle.D_NOT_ON_CURVE and oe.value.index == 6


def test_raw_contribute_shape_2p20(ctx, oracle):
    """The documented shim of phase-2 batch_exp at config-3 size: 2^20 raw points, one scalar, in place."""
    n = 1 << 20
    gen = np.frombuffer((1).to_bytes(32, "big") + (2).to_bytes(32, "big"), dtype=np.uint8)
    tau = be(0x1d7a3f6c2b9e80415f6a7b8c9d0e1f2031425364758697a8b9cadbecfd0e1f21 % R_MOD)
    raw = ctx.batch_mul_powers(0, np.tile(gen, n), tau, None, 1, 0, RAW)
    wire = ctx.recode(0, raw, RAW, 0)
    k = be(0x0fedcba987654321fedcba987654321fedcba987654321fedcba987654321 % R_MOD)
    buf = raw.copy()
    got = ctx.batch_mul(0, buf, k, RAW, RAW, out=buf)                     # in == out, as the shim calls it
    exp = oracle.batch_mul(0, wire.tobytes(), k, threads=32)
    assert ctx.recode(0, got, RAW, 0).tobytes() == exp
    m = 4096                                                              # and the raw bytes themselves on a prefix
    assert got[: 64 * m].tobytes() == wire_to_raw(exp[: 64 * m])


def test_raw_g2_layout(ctx, oracle):
    """include/p2b.h defines the same layout for G2 (x.c0, x.c1, y.c0, y.c1 Montgomery limbs); used internally by the group
    FFT and sparse paths.  The reference implements RawEncodable for G1 only."""
    n = 50
    pts = random_points(oracle, 1, n, seed=951)
    raw = ctx.recode(1, pts, 0, RAW).tobytes()
    assert raw == oracle.batch_mul(1, pts, be(1), 0, RAW, threads=4)
    k = be(987654321987654321)
    assert ctx.batch_mul(1, raw, k, RAW, 0).tobytes() == oracle.batch_mul(1, pts, k, threads=4)
