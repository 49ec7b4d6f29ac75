"""GPU parity: Pippenger MSM through the C ABI vs the oracle's restatement of bellman's multiexp (bit-exact)."""
import numpy as np
import pytest

from util import EDGE_SCALARS, R_MOD, be, random_points, random_scalars

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("group", [0, 1])
@pytest.mark.parametrize("n", [1, 2, 31, 100, 1000, 5000])
def test_msm_matches_oracle(ctx, oracle, group, n):
    pts = random_points(oracle, group, n, seed=300 + n)
    sc = random_scalars(n, seed=400 + n)
    assert ctx.msm(group, pts, sc) == oracle.msm(group, pts, sc, threads=8)


@pytest.mark.parametrize("group", [0, 1])
def test_msm_edge_cases(ctx, oracle, group):
    size = 128 if group else 64
    n = len(EDGE_SCALARS)
    pts = bytearray(random_points(oracle, group, n, seed=41))
    sc = b"".join(be(k) for k in EDGE_SCALARS)
    assert ctx.msm(group, bytes(pts), sc) == oracle.msm(group, bytes(pts), sc, threads=4)
    # repeated points (P + P hits the doubling path), P and -P with equal scalars (cancels), infinity inputs
    p0 = bytes(pts[:size])
    neg = oracle.batch_mul(group, p0, be(R_MOD - 1))
    inf = bytes([0x40]) + bytes(size - 1)
    pts2 = p0 * 40 + neg * 3 + inf * 2 + bytes(pts[size:3 * size])
    n2 = len(pts2) // size
    sc2 = be(5) * 40 + be(5) * 3 + be(77) * 2 + random_scalars(2, 9)
    assert ctx.msm(group, pts2, sc2) == oracle.msm(group, pts2, sc2, threads=4)
    # everything cancels -> infinity
    assert ctx.msm(group, p0 + neg, be(12345) * 2) == inf
    # all-zero scalars / empty input -> infinity
    assert ctx.msm(group, bytes(pts), bytes(32 * n)) == inf
    assert ctx.msm(group, b"", b"") == inf


def test_msm_large_linearity(ctx, oracle):
    """2^18 terms: MSM(k) + MSM(k') == MSM(k + k') (size-independent check) and shard-sum == whole
    (the multi-GPU decomposition: per-range results combined with sum_points)."""
    n = 1 << 18
    base = np.frombuffer(random_points(oracle, 0, 512, seed=51), dtype=np.uint8)
    pts = np.tile(base, n // 512)
    rng = np.random.default_rng(52)
    k1 = rng.integers(0, 256, size=(n, 32), dtype=np.uint8); k1[:, 0] &= 0x1f
    k2 = rng.integers(0, 256, size=(n, 32), dtype=np.uint8); k2[:, 0] &= 0x0f
    ks = (np.array([int.from_bytes(a.tobytes(), "big") for a in k1[:64]], dtype=object))
    a = ctx.msm(0, pts, k1.reshape(-1))
    b = ctx.msm(0, pts, k2.reshape(-1))
    ksum = np.frombuffer(b"".join(be((int.from_bytes(x.tobytes(), "big") + int.from_bytes(y.tobytes(), "big")) % R_MOD)
                                  for x, y in zip(k1, k2)), dtype=np.uint8)
    c = ctx.msm(0, pts, ksum)
    assert ctx.sum_points(0, a + b) == c
    parts = b"".join(ctx.msm(0, pts[s * (n // 4) * 64:(s + 1) * (n // 4) * 64], k1.reshape(-1)[s * (n // 4) * 32:(s + 1) * (n // 4) * 32])
                     for s in range(4))
    assert ctx.sum_points(0, parts) == a
    # spot check against the oracle on a prefix
    m = 3000
    assert ctx.msm(0, pts[:m * 64], k1.reshape(-1)[:m * 32]) == oracle.msm(0, pts[:m * 64].tobytes(), k1.reshape(-1)[:m * 32].tobytes(), threads=8)


def test_msm_rejects_bad_input(ctx, oracle):
    from phase2_bn254_b200 import lib
    pts = bytearray(random_points(oracle, 0, 8, seed=61))
    pts[64 * 3] |= 0x80
    with pytest.raises(lib.P2BError) as e:
        ctx.msm(0, bytes(pts), random_scalars(8, 1))
    assert e.value.code == lib.EDECODE and e.value.index == 3
    with pytest.raises(lib.P2BError) as e:
        ctx.msm(0, random_points(oracle, 0, 2, seed=62), be(1) + b"\xff" * 32)
    assert e.value.code == lib.EARG and e.value.index == 1


@pytest.mark.parametrize("group", [0, 1])
def test_msm_streamed_host_path(ctx, oracle, group, monkeypatch):
    """Host buffers above the streaming threshold are processed chunk by chunk into the same buckets (H2D overlapped
    with the sort / accumulate of the previous chunk); the threshold is lowered through the test hook."""
    n = 3000
    pts = bytearray(random_points(oracle, group, n, seed=71))
    size = 128 if group else 64
    pts[size * 1500: size * 1501] = bytes([0x40]) + bytes(size - 1)
    sc = random_scalars(n, seed=72)
    exp = oracle.msm(group, bytes(pts), sc, threads=8)
    for chunk in (700, 1000, 2999):
        monkeypatch.setenv("P2B_MSM_STREAM_CHUNK", str(chunk))
        assert ctx.msm(group, bytes(pts), sc) == exp
    monkeypatch.setenv("P2B_MSM_STREAM_CHUNK", "512")
    from phase2_bn254_b200 import lib
    bad = bytearray(pts)
    bad[size * 2000] |= 0x80
    with pytest.raises(lib.P2BError) as e:
        ctx.msm(group, bytes(bad), sc)
    assert e.value.code == lib.EDECODE and e.value.index == 2000


@pytest.mark.parametrize("group", [0, 1])
def test_power_pairs(ctx, oracle, group):
    """The verifier's random linear combination (utils.rs:112-135) as two MSMs; same_ratio then holds for a geometric
    progression: b == [tau] a."""
    from phase2_bn254_b200.powersoftau import power_pairs
    from util import G1_GEN, G2_GEN
    n = 257
    tau = be(0x1234567 ** 9 % R_MOD)
    gen = G2_GEN if group else G1_GEN
    v = oracle.batch_mul_powers(group, gen * n, tau, None, 0, threads=8)     # tau^i G
    sc = random_scalars(n - 1, seed=91)
    a, b = power_pairs(ctx, group, v, sc)
    size = 128 if group else 64
    assert a == oracle.msm(group, v[: (n - 1) * size], sc, threads=8)
    assert b == oracle.msm(group, v[size:], sc, threads=8)
    assert b == oracle.point_mul(group, a, tau)


@pytest.mark.parametrize("group", [0, 1])
def test_msm_skewed_scalars(ctx, oracle, group, monkeypatch):
    """Equal scalars put every term of a window into ONE bucket: the per-thread segment cap hands the tail to the
    block-level heavy-bucket path.  Also forced with tiny segments on random scalars (several items per bucket)."""
    n = 6000 if group == 0 else 2500
    pts = random_points(oracle, group, n, seed=95)
    k = be(0x1234567890abcdef1234567890abcdef1234567890abcdef1234567 % R_MOD)
    assert ctx.msm(group, pts, k * n) == oracle.msm(group, pts, k * n, threads=8)
    mixed = k * (n // 2) + random_scalars(n - n // 2, seed=96)
    assert ctx.msm(group, pts, mixed) == oracle.msm(group, pts, mixed, threads=8)
    monkeypatch.setenv("P2B_MSM_SEG", "3")
    sc = random_scalars(n, seed=97)
    exp = oracle.msm(group, pts, sc, threads=8)
    assert ctx.msm(group, pts, sc) == exp
    monkeypatch.setenv("P2B_MSM_STREAM_CHUNK", "1000")          # streamed chunks continue the same buckets
    assert ctx.msm(group, pts, sc) == exp
    assert ctx.msm(group, pts, k * n) == oracle.msm(group, pts, k * n, threads=8)


@pytest.mark.parametrize("group", [0, 1])
def test_msm_streamed_copy_bound_plan(ctx, oracle, group, monkeypatch):
    """The chunk plan used when the host link is the slower side (small chunks at the end, P2B_MSM_TAIL forces it): same sum."""
    n = 3000
    pts = random_points(oracle, group, n, seed=171)
    sc = random_scalars(n, seed=172)
    exp = oracle.msm(group, pts, sc, threads=8)
    for tail in ("1", "0"):
        monkeypatch.setenv("P2B_MSM_TAIL", tail)
        for chunk in (512, 1000, 2999):
            monkeypatch.setenv("P2B_MSM_STREAM_CHUNK", str(chunk))
            assert ctx.msm(group, pts, sc) == exp, (tail, chunk)
