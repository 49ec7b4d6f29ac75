"""GPU parity: the G2 batch subgroup probe (csrc/msm_g2.cu `g2_subgroup_probe`, include/p2b.h P2B_G2_EXACT).

The reference multiplies whatever curve point it decoded (no subgroup check for BN256 G2, pairing/src/bn256/ec.rs:1145-1213),
so the product must return the exact [k]P for EVERY point of E'(Fq2).  Large G2 batches are first proven to lie in the order-r
subgroup (then the endomorphism split is valid) or else take the exact path; these tests force the probe at small sizes
(P2B_G2_PROBE_MIN) and compare with the oracle for honest batches, batches hiding a point outside the subgroup (large and
smallest possible order of the cofactor component), off-curve garbage, infinity, compressed input and the tau-powers shape."""
import numpy as np
import pytest

from util import EDGE_SCALARS, R_MOD, be, random_points, random_scalars, twist_points_outside_subgroup

pytestmark = pytest.mark.gpu


@pytest.fixture()
def probe_small(monkeypatch):
    monkeypatch.setenv("P2B_G2_PROBE_MIN", "64")


def test_honest_batch_is_proven_and_takes_the_split_path(ctx, oracle, probe_small):
    n = len(EDGE_SCALARS) + 400
    pts = random_points(oracle, 1, n, seed=41)
    sc = b"".join(be(k) for k in EDGE_SCALARS) + random_scalars(400, seed=42)
    before, _ = ctx.g2_probe_stats()
    got = ctx.batch_mul(1, pts, sc, 0, 1).tobytes()
    after, verdict = ctx.g2_probe_stats()
    assert after == before + 1 and verdict == 0
    assert got == oracle.batch_mul(1, pts, sc, 0, 1, threads=8)
    # broadcast scalar and the tau-powers shape, with a point at infinity in the batch
    buf = bytearray(pts)
    buf[17 * 128:18 * 128] = bytes([0x40]) + bytes(127)
    k = be(R_MOD - 0x1234567)
    assert ctx.batch_mul(1, bytes(buf), k).tobytes() == oracle.batch_mul(1, bytes(buf), k, threads=8)
    assert ctx.g2_probe_stats() == (after + 1, 0)
    tau, coeff = be(0x2222 ** 15 % R_MOD), be(R_MOD - 77)
    got = ctx.batch_mul_powers(1, pts, tau, coeff, (1 << 21) + 5, 0, 1).tobytes()
    assert got == oracle.batch_mul_powers(1, pts, tau, coeff, (1 << 21) + 5, 0, 1, threads=8)
    assert ctx.g2_probe_stats() == (after + 2, 0)


@pytest.mark.parametrize("small_order", [False, True])
def test_point_outside_the_subgroup_sends_the_batch_down_the_exact_path(ctx, oracle, probe_small, small_order):
    n = 300
    buf = bytearray(random_points(oracle, 1, n, seed=43))
    bad = twist_points_outside_subgroup(2, seed=44 + small_order, small_order=small_order)
    buf[123 * 128:124 * 128] = bad[:128]
    buf[299 * 128:300 * 128] = bad[128:]
    sc = random_scalars(n, seed=45)
    exp = oracle.batch_mul(1, bytes(buf), sc, 0, 0, threads=8)
    for trial in range(3):                       # fresh coefficients every time
        before, _ = ctx.g2_probe_stats()
        got = ctx.batch_mul(1, bytes(buf), sc, 0, 0).tobytes()
        assert ctx.g2_probe_stats() == (before + 1, 1)
        assert got == exp
    # the exact flag gives the same bytes without a probe; the caller's (false) promise is the caller's problem and is not tested
    from phase2_bn254_b200 import lib
    before, _ = ctx.g2_probe_stats()
    assert ctx.batch_mul(1, bytes(buf), sc, 0, 0, flags=lib.G2_EXACT).tobytes() == exp
    assert ctx.g2_probe_stats()[0] == before


def test_off_curve_point_is_caught_deterministically(ctx, oracle, probe_small):
    n = 200
    buf = bytearray(random_points(oracle, 1, n, seed=46))
    buf[77 * 128 + 127] ^= 1                     # y.c0 changed: not on the curve (unchecked mode computes on it, like the reference)
    sc = random_scalars(n, seed=47)
    before, _ = ctx.g2_probe_stats()
    got = ctx.batch_mul(1, bytes(buf), sc).tobytes()
    assert ctx.g2_probe_stats() == (before + 1, 1)
    assert got == oracle.batch_mul(1, bytes(buf), sc, threads=8)
    from phase2_bn254_b200 import lib
    with pytest.raises(lib.P2BError) as e:
        ctx.batch_mul(1, bytes(buf), sc, flags=lib.CHECK_INPUT)
    assert e.value.code == lib.EDECODE and e.value.sub == lib.DEC_NOT_ON_CURVE and e.value.index == 77


def test_compressed_input_and_decode_errors(ctx, oracle, probe_small):
    from phase2_bn254_b200 import lib
    n = 260
    pts = random_points(oracle, 1, n, seed=48)
    sc = random_scalars(n, seed=49)
    comp = oracle.batch_mul(1, pts, be(1), 0, 1, threads=8)
    before, _ = ctx.g2_probe_stats()
    assert ctx.batch_mul(1, comp, sc, 1, 1).tobytes() == oracle.batch_mul(1, comp, sc, 1, 1, threads=8)
    assert ctx.g2_probe_stats() == (before + 1, 0)
    bad = bytearray(pts)
    bad[5 * 128] |= 0x80
    with pytest.raises(lib.P2BError) as e:
        ctx.batch_mul(1, bytes(bad), sc)
    assert e.value.code == lib.EDECODE and e.value.sub == lib.DEC_UNEXPECTED_COMPRESSION_MODE and e.value.index == 5
    bad = bytearray(pts)
    bad[9 * 128:10 * 128] = bytes([0x40]) + bytes(127)
    with pytest.raises(lib.P2BError) as e:
        ctx.batch_mul(1, bytes(bad), sc, flags=lib.REJECT_INFINITY)
    assert e.value.code == lib.EINFINITY_IN and e.value.index == 9


def test_default_threshold_2p17_points(ctx, oracle):
    """The probe as callers get it (no test hook): 2^17 G2 points.  Honest batch: proven, split path, sample vs the oracle and
    [a]([b]P) == [ab]P on all points; with one point outside the subgroup hidden in it: exact path, that point vs the oracle."""
    n = 1 << 17
    base = np.frombuffer(random_points(oracle, 1, 512, seed=50), dtype=np.uint8)
    pts = np.tile(base, n // 512)
    a, b = 0x0fedcba987654321 ** 3 % R_MOD, R_MOD - 424242
    before, _ = ctx.g2_probe_stats()
    p1 = ctx.batch_mul(1, pts, be(a))
    assert ctx.g2_probe_stats() == (before + 1, 0)
    p2 = ctx.batch_mul(1, p1, be(b))
    p3 = ctx.batch_mul(1, pts, be(a * b % R_MOD))
    assert ctx.g2_probe_stats() == (before + 3, 0)
    assert np.array_equal(p2, p3)
    assert p3[:64 * 128].tobytes() == oracle.batch_mul(1, pts[:64 * 128].tobytes(), be(a * b % R_MOD), threads=8)
    hidden = np.array(pts)
    bad = twist_points_outside_subgroup(1, seed=51, small_order=True)
    at = 98765
    hidden[at * 128:(at + 1) * 128] = np.frombuffer(bad, dtype=np.uint8)
    got = ctx.batch_mul(1, hidden, be(a))
    assert ctx.g2_probe_stats() == (before + 4, 1)
    assert got[at * 128:(at + 1) * 128].tobytes() == oracle.batch_mul(1, bad, be(a), threads=1)
    assert np.array_equal(got[:at * 128], p1[:at * 128]) and np.array_equal(got[(at + 1) * 128:], p1[(at + 1) * 128:])


def test_transform_and_group_fft_with_probed_g2_sections(ctx, oracle, probe_small):
    """`transform` (TauG2 section, tau-powers shape inside the chunk pipeline, compressed and uncompressed input) and a G2 group
    FFT with the probe active at these sizes: the response / the transform equal the oracle's."""
    from phase2_bn254_b200.powersoftau import BatchedAccumulator, CeremonyParams, PrivateKey
    tau, alpha, beta = 0x1111 ** 17 % R_MOD, 0x2222 ** 13 % R_MOD, 0x3333 ** 11 % R_MOD
    size, batch = 8, 64
    params = CeremonyParams(size, batch)
    ch0 = oracle.pot_generate_initial(size)
    ch1c = oracle.pot_transform(ch0, size, batch, be(tau), be(alpha), be(beta), out_compressed=True, threads=8)
    ch1 = oracle.pot_transform(ch0, size, batch, be(tau), be(alpha), be(beta), out_compressed=False, threads=8)
    key = PrivateKey(alpha, beta, tau)
    for in_c, src in ((False, ch1), (True, ch1c)):
        exp = oracle.pot_transform(src, size, batch, be(alpha), be(beta), be(tau), in_c, True, True, threads=8)
        out = np.zeros(len(exp), dtype=np.uint8)
        before, _ = ctx.g2_probe_stats()
        BatchedAccumulator.transform(np.frombuffer(src, dtype=np.uint8), out, in_c, True, True, key, params, ctx=ctx)
        after, verdict = ctx.g2_probe_stats()
        assert after > before and verdict == 0
        assert out[64:].tobytes() == exp[64:], in_c
    # one G2 point of the challenge replaced by a point outside the subgroup: still the reference's bytes (exact path)
    bad = twist_points_outside_subgroup(1, seed=52, small_order=True)
    buf = bytearray(ch1)
    off = 64 + params.powers_g1_length * 64 + 37 * 128
    buf[off:off + 128] = bad
    exp = oracle.pot_transform(bytes(buf), size, batch, be(alpha), be(beta), be(tau), False, True, True, threads=8)
    out = np.zeros(len(exp), dtype=np.uint8)
    BatchedAccumulator.transform(np.frombuffer(bytes(buf), dtype=np.uint8), out, False, True, True, key, params, ctx=ctx)
    assert out[64:].tobytes() == exp[64:]
    d = 256
    pts = random_points(oracle, 1, d, seed=53)
    got = ctx.group_fft(1, pts, True)
    exp = ctx.group_fft(1, pts, True, flags=8)                 # P2B_G2_EXACT: the path the earlier rounds checked against the definition
    assert np.array_equal(got, exp)
