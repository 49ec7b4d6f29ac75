"""GPU parity: FFT over group elements (bellman EvaluationDomain over Point<G>, domain.rs:154-174 + group.rs:30-50) and the
prepare_phase2 step (powersoftau/src/bin/prepare_phase2.rs:62-241).  Checked against (i) the DFT definition evaluated
with the oracle's MSM, (ii) the closed form for a geometric progression tau^i G, whose inverse transform is L_j(tau) G
with the Lagrange basis evaluated in plain python integers, (iii) round trips."""
import numpy as np
import pytest

from util import G1_GEN, G2_GEN, R_MOD, be, random_points

pytestmark = pytest.mark.gpu

ROOT_OF_UNITY = pow(7, (R_MOD - 1) >> 28, R_MOD)
TAU = 0x1111111111111111111111111111111111111111111111111111111111111111 % R_MOD
ALPHA = 0x2222222222222222222222222222222222222222222222222222222222222222 % R_MOD
BETA = 0x0333333333333333333333333333333333333333333333333333333333333333 % R_MOD


def omega(log_d):
    return pow(ROOT_OF_UNITY, 1 << (28 - log_d), R_MOD)


def lagrange_at(tau, log_d):
    """L_j(tau) = (1/d) sum_i omega^(-ij) tau^i for j < d."""
    d = 1 << log_d
    wi = pow(omega(log_d), -1, R_MOD)
    dinv = pow(d, -1, R_MOD)
    pw = [pow(tau, i, R_MOD) for i in range(d)]
    return [sum(pow(wi, i * j, R_MOD) * pw[i] for i in range(d)) * dinv % R_MOD for j in range(d)]


@pytest.mark.parametrize("group", [0, 1])
@pytest.mark.parametrize("log_d", [0, 1, 3])
def test_group_fft_vs_definition(ctx, oracle, group, log_d):
    d = 1 << log_d
    size = 128 if group else 64
    pts = bytearray(random_points(oracle, group, d, seed=600 + log_d))
    if d >= 4:
        pts[size * 2: size * 3] = bytes([0x40]) + bytes(size - 1)          # infinity is a valid group element
        pts[size * 3: size * 4] = pts[0:size]                              # equal points (doubling inside the butterfly)
    pts = bytes(pts)
    w = omega(log_d)
    fwd = b"".join(oracle.msm(group, pts, b"".join(be(pow(w, i * j, R_MOD)) for i in range(d)), threads=4) for j in range(d))
    wi, dinv = pow(w, -1, R_MOD), pow(d, -1, R_MOD)
    inv = b"".join(oracle.msm(group, pts, b"".join(be(pow(wi, i * j, R_MOD) * dinv % R_MOD) for i in range(d)), threads=4)
                   for j in range(d))
    assert ctx.group_fft(group, pts, False).tobytes() == fwd
    assert ctx.group_fft(group, pts, True).tobytes() == inv
    comp = oracle.batch_mul(group, pts, be(1), 0, 1)
    assert ctx.group_fft(group, comp, True, 1, 1).tobytes() == oracle.batch_mul(group, inv, be(1), 0, 1)


@pytest.mark.parametrize("group,log_d", [(0, 6), (0, 10), (1, 7)])
def test_group_ifft_of_powers_is_lagrange_basis(ctx, oracle, group, log_d):
    d = 1 << log_d
    gen = G2_GEN if group else G1_GEN
    powers = oracle.batch_mul_powers(group, gen * d, be(TAU), None, 0, threads=8)             # tau^i G
    exp = oracle.batch_mul(group, gen * d, b"".join(be(v) for v in lagrange_at(TAU, log_d)), threads=8)
    got = ctx.group_fft(group, powers, True)
    assert got.tobytes() == exp
    assert ctx.group_fft(group, got, False).tobytes() == powers                               # fft(ifft(x)) == x


def test_group_fft_round_trip_large(ctx, oracle):
    log_d = 13
    d = 1 << log_d
    base = np.frombuffer(random_points(oracle, 0, 256, seed=77), dtype=np.uint8)
    pts = np.tile(base, d // 256)
    f = ctx.group_fft(0, pts, False)
    assert np.array_equal(ctx.group_fft(0, f, True), pts)


@pytest.mark.parametrize("compressed", [False, True])
def test_prepare_phase2_matches_closed_form(ctx, oracle, compressed):
    from phase2_bn254_b200.powersoftau import CeremonyParams, prepare_phase2
    size, batch = 5, 8
    params = CeremonyParams(size, batch)
    ch0 = oracle.pot_generate_initial(size)
    acc = oracle.pot_transform(ch0, size, batch, be(TAU), be(ALPHA), be(BETA), out_compressed=compressed, threads=8)
    for m in range(size + 1):
        d = 1 << m
        lag = lagrange_at(TAU, m)
        g1 = lambda ks: oracle.batch_mul(0, G1_GEN * len(ks), b"".join(be(k % R_MOD) for k in ks), threads=8) if ks else b""
        g2 = lambda ks: oracle.batch_mul(1, G2_GEN * len(ks), b"".join(be(k % R_MOD) for k in ks), threads=8)
        exp = (g1([ALPHA]) + g1([BETA]) + g2([BETA]) + g1(lag) + g2(lag) + g1([ALPHA * l for l in lag]) +
               g1([BETA * l for l in lag]) + g1([pow(TAU, i + d, R_MOD) - pow(TAU, i, R_MOD) for i in range(d - 1)]))
        got = prepare_phase2(ctx, np.frombuffer(acc, dtype=np.uint8), params, m, input_is_compressed=compressed)
        assert len(exp) == 192 + 384 * d == got.size
        assert got.tobytes() == exp, "m = %d" % m


def test_prepare_phase2_rejects_bad_input(ctx, oracle):
    from phase2_bn254_b200.powersoftau import CeremonyParams, DeserializationError, prepare_phase2
    size = 3
    params = CeremonyParams(size, 4)
    acc = bytearray(oracle.pot_transform(oracle.pot_generate_initial(size), size, 4, be(TAU), be(ALPHA), be(BETA),
                                         out_compressed=False, threads=2))
    acc[64 + 64 * 2 + 63] ^= 1                                                               # tau_g1[2] off the curve
    with pytest.raises(DeserializationError):
        prepare_phase2(ctx, np.frombuffer(bytes(acc), dtype=np.uint8), params, 2, input_is_compressed=False)


@pytest.mark.parametrize("group,log_d", [(0, 6), (1, 5)])
@pytest.mark.parametrize("world", [2, 4, 8])
def test_sharded_group_fft_equals_single_gpu(ctx, oracle, group, log_d, world):
    """The multi-GPU schedule (rank-crossing DIF stages through p2b_g*_gfft_stage, block-local transforms through
    p2b_g*_group_fft_scaled, outputs at X[R j + bitrev(r)]) with all ranks simulated on one GPU: the bytes of the ordinary
    single-GPU transform, both directions, with infinity and repeated points in the input."""
    from phase2_bn254_b200.dist import simulate_sharded_group_fft
    d = 1 << log_d
    size = 128 if group else 64
    pts = bytearray(random_points(oracle, group, d, seed=650 + log_d))
    pts[size * 2: size * 3] = bytes([0x40]) + bytes(size - 1)
    pts[size * 3: size * 4] = pts[0:size]
    pts[size * (d // 2): size * (d // 2 + 1)] = pts[0:size]                # a == b across the first rank-crossing stage
    pts = bytes(pts)
    for inverse in (False, True):
        exp = ctx.group_fft(group, pts, inverse).tobytes()
        assert simulate_sharded_group_fft(ctx, group, pts, world, inverse).tobytes() == exp, (world, inverse)


def test_gfft_stage_matches_oracle(ctx, oracle):
    """One rank-crossing stage by itself: a + b and [w^(start+i)](a - b) against the oracle's point arithmetic."""
    from util import OracleCtx
    n = 16
    for group in (0, 1):
        a, b = random_points(oracle, group, n, seed=660 + group), random_points(oracle, group, n, seed=670 + group)
        w = np.frombuffer(be(omega(7)), dtype=np.uint8)
        s, dd = ctx.gfft_stage(group, a, b, w, 37)
        es, ed = OracleCtx(oracle).gfft_stage(group, a, b, w, 37)
        assert s.tobytes() == es.tobytes() and dd.tobytes() == ed.tobytes()
        _, plain = ctx.gfft_stage(group, a, b, None, want_sum=False)          # w = NULL: plain differences (the H query)
        assert plain.tobytes() == OracleCtx(oracle).gfft_stage(group, a, b, None, want_sum=False)[1].tobytes()


@pytest.mark.parametrize("compressed", [False, True])
def test_sharded_prepare_phase2_assembly(ctx, oracle, compressed):
    """dist.sharded_prepare_phase2 on one rank (no exchange) writes the bytes of p2b_pot_prepare_phase2: section offsets, the
    strided output placement and the H query's index ranges."""
    from phase2_bn254_b200.dist import sharded_prepare_phase2
    from phase2_bn254_b200.powersoftau import CeremonyParams
    size = 4
    prm = CeremonyParams(size, 8)
    ch0 = oracle.pot_generate_initial(size)
    acc = oracle.pot_transform(ch0, size, 8, be(TAU), be(ALPHA), be(BETA), out_compressed=compressed, threads=8)
    amap = np.frombuffer(acc, dtype=np.uint8)
    for m in (1, 3, 4):
        exp = ctx.pot_prepare_phase2(amap, size, m, compressed_input=compressed)
        out = np.zeros(exp.size, dtype=np.uint8)
        sharded_prepare_phase2(ctx, amap, prm, m, out, 0, 1, input_is_compressed=compressed)
        assert out.tobytes() == exp.tobytes(), m
