"""CPU: the C oracle against the independent python big-int restatement (oracle/bn254_ref.py) on seeded random
and edge inputs, codec error behaviour (pairing/src/bn256/ec.rs:772-826,875-919,1145-1213,1264-1315), and the GLV
constants the CUDA path relies on."""
import random

import pytest

import bn254_ref as ref
from util import EDGE_SCALARS, G1_GEN, G2_GEN, LAMBDA, Q_MOD, R_MOD, be, random_points, random_scalars


def test_field_ops_vs_bigint(oracle):
    rng = random.Random(1)
    for field, mod in ((0, Q_MOD), (1, R_MOD)):
        vals = [0, 1, 2, mod - 1, mod - 2, (mod - 1) // 2] + [rng.randrange(mod) for _ in range(200)]
        for i in range(0, len(vals) - 1):
            a, b = vals[i], vals[i + 1]
            assert oracle.field_op(field, 0, be(a), be(b)) == be(a * b % mod)
            assert oracle.field_op(field, 1, be(a), be(b)) == be((a + b) % mod)
            assert oracle.field_op(field, 2, be(a), be(b)) == be((a - b) % mod)
            inv = oracle.field_op(field, 3, be(a))
            assert (inv is None) if a == 0 else inv == be(pow(a, -1, mod))
        assert oracle.field_op(field, 0, be(mod), be(1)) is None          # not in field
    for _ in range(50):
        a = rng.randrange(Q_MOD)
        s = oracle.field_op(0, 4, be(a))
        if s is None:
            assert pow(a, (Q_MOD - 1) // 2, Q_MOD) == Q_MOD - 1
        else:
            assert pow(int.from_bytes(s, "big"), 2, Q_MOD) == a


@pytest.mark.parametrize("group", [0, 1])
def test_point_mul_vs_bigint(oracle, group):
    curve, gen, enc, dec = ((ref.G1, ref.G1_GEN, ref.g1_encode, ref.g1_decode) if group == 0 else
                            (ref.G2, ref.G2_GEN, ref.g2_encode, ref.g2_decode))
    rng = random.Random(2 + group)
    ks = EDGE_SCALARS + [rng.randrange(R_MOD) for _ in range(6)]
    pts = random_points(oracle, group, len(ks), seed=3)
    size = 128 if group else 64
    exp_u = b"".join(enc(curve.mul(dec(pts[i * size:(i + 1) * size], False), k), False) for i, k in enumerate(ks))
    exp_c = b"".join(enc(curve.mul(dec(pts[i * size:(i + 1) * size], False), k), True) for i, k in enumerate(ks))
    sc = b"".join(be(k) for k in ks)
    assert oracle.batch_mul(group, pts, sc, threads=3) == exp_u
    assert oracle.batch_mul(group, pts, sc, 0, 1, threads=1) == exp_c
    assert oracle.batch_mul(group, exp_c, be(1), 1, 0) == exp_u            # decompress
    for i, k in enumerate(ks[:8]):
        assert oracle.point_mul(group, pts[i * size:(i + 1) * size], be(k)) == exp_u[i * size:(i + 1) * size]


@pytest.mark.parametrize("group", [0, 1])
def test_codec_errors(oracle, group):
    size = 128 if group else 64
    p = bytearray(random_points(oracle, group, 1, seed=5))
    # not on curve
    bad = bytearray(p); bad[size - 1] ^= 1
    with pytest.raises(oracle.OracleError) as e:
        oracle.point_recode(group, bytes(bad), 0, 0, True)
    assert e.value.sub == oracle.D_NOT_ON_CURVE
    assert oracle.point_recode(group, bytes(bad), 0, 0, False) == bytes(bad)   # unchecked accepts
    # coordinate >= q
    bad = bytearray(p); bad[0:32] = be(Q_MOD)
    with pytest.raises(oracle.OracleError) as e:
        oracle.point_recode(group, bytes(bad), 0, 0, False)
    assert e.value.sub == oracle.D_COORD
    # sign bit on an uncompressed encoding: G1 UnexpectedInformation, G2 UnexpectedCompressionMode
    bad = bytearray(p); bad[0] |= 0x80
    with pytest.raises(oracle.OracleError) as e:
        oracle.point_recode(group, bytes(bad), 0, 0, False)
    assert e.value.sub == (oracle.D_UNEXPECTED_COMPRESSION if group else oracle.D_UNEXPECTED_INFO)
    # infinity flag with stray bits
    bad = bytearray(size); bad[0] = 0x40; bad[size - 1] = 1
    with pytest.raises(oracle.OracleError) as e:
        oracle.point_recode(group, bytes(bad), 0, 0, False)
    assert e.value.sub == oracle.D_UNEXPECTED_INFO
    inf = bytes([0x40]) + bytes(size - 1)
    assert oracle.point_recode(group, inf, 0, 1, True) == bytes([0x40]) + bytes(size // 2 - 1)
    # phase-1 semantics: infinity in the input is an error, phase-2 tolerates it
    with pytest.raises(oracle.OracleError) as e:
        oracle.batch_mul(group, bytes(p) + inf, be(3), reject_inf=True)
    assert e.value.code == oracle.EINFINITY_IN and e.value.index == 1
    assert oracle.batch_mul(group, bytes(p) + inf, be(3))[size:] == inf
    # k = 0 / k = r produce infinity: phase-1 rejects the output
    with pytest.raises(oracle.OracleError) as e:
        oracle.batch_mul(group, bytes(p), be(0), reject_inf=True)
    assert e.value.code == oracle.EINFINITY_OUT


def test_transform_vs_bigint_and_thread_invariance(oracle):
    size, batch = 2, 3
    rng = random.Random(7)
    tau, alpha, beta = (rng.randrange(1, R_MOD) for _ in range(3))
    ch = oracle.pot_generate_initial(size)
    params = ref.CeremonyParams(size, batch)
    assert ch == ref.generate_initial(params)
    assert len(ch) == params.accumulator_size == oracle.accumulator_size(size, False)
    exp = ref.transform(params, ch, tau, alpha, beta)
    for threads in (1, 2, 5):
        got = oracle.pot_transform(ch, size, batch, be(tau), be(alpha), be(beta), threads=threads)
        assert got[64:] == exp
    assert len(got) + 768 == params.contribution_size


def test_msm_and_fft_vs_bigint(oracle):
    n = 40
    pts = random_points(oracle, 0, n, seed=11)
    sc = random_scalars(n, seed=12)
    ps = [ref.g1_decode(pts[i * 64:(i + 1) * 64], False) for i in range(n)]
    ks = [int.from_bytes(sc[i * 32:(i + 1) * 32], "big") for i in range(n)]
    assert oracle.msm(0, pts, sc, threads=3) == ref.g1_encode(ref.msm(ref.G1, ps, ks), False)
    rng = random.Random(13)
    for log_n in (2, 6, 9):
        a = [rng.randrange(R_MOD) for _ in range(1 << log_n)]
        data = b"".join(be(x) for x in a)
        for inv, cos in ((0, 0), (1, 0), (0, 1), (1, 1)):
            exp = b"".join(be(x) for x in ref.fft(a, bool(inv), bool(cos)))
            assert oracle.fr_fft(data, inv, cos, threads=1) == exp
            assert oracle.fr_fft(data, inv, cos, threads=8) == exp          # parallel_fft == serial_fft (domain.rs:515)
        assert oracle.fr_fft(oracle.fr_fft(data), True) == data               # fft_composition (domain.rs:428)


def test_glv_constants():
    """phi(x, y) = (beta x, y) = [lambda](x, y) for the constants stored in csrc/smul.cuh (G1_BETA is Montgomery)."""
    limbs = [0xd782e155, 0x71930c11, 0xffbe3323, 0xa6bb947c, 0xd4741444, 0xaa303344, 0x26594943, 0x2c3b3f0d]
    beta = sum(v << (32 * i) for i, v in enumerate(limbs)) * pow(2**256, -1, Q_MOD) % Q_MOD
    assert beta != 1 and pow(beta, 3, Q_MOD) == 1
    assert (LAMBDA * LAMBDA + LAMBDA + 1) % R_MOD == 0
    for k in (1, 5, 0x1234567):
        p = ref.G1.mul(ref.G1_GEN, k)
        assert ref.G1.mul(p, LAMBDA) == (beta * p[0] % Q_MOD, p[1])
    # lattice basis of the decomposition: a + b * lambda = 0 mod r
    a1, b1 = 0x89d3256894d213e3, -0x6f4d8248eeb859fc8211bbeb7d4f1128
    a2, b2 = 0x6f4d8248eeb859fd0be4e1541221250b, 0x89d3256894d213e3
    assert (a1 + b1 * LAMBDA) % R_MOD == 0 and (a2 + b2 * LAMBDA) % R_MOD == 0


def test_g2_cofactor_structure_behind_the_subgroup_probe():
    """The facts csrc/msm_g2.cu `g2_subgroup_probe` rests on, checked with the big-int reference: #E'(Fq2) = r h with
    h = 2q - r (ec.rs:1347-1357) = 10069 * 5864401 * 1875725156269 * p177 and gcd(r, h) = 1, so [r]W = O exactly for the
    points of the order-r subgroup; a point with a cofactor component of the smallest possible order (10069) is on the
    curve, is not killed by r, and a random combination hiding it is not killed either unless its coefficient is a multiple
    of 10069; sums of subgroup points stay in the subgroup."""
    import math
    import random

    from util import G2_COFACTOR, G2_COFACTOR_SMALL_PRIME, twist_points_outside_subgroup
    assert G2_COFACTOR == 0x30644e72e131a029b85045b68181585e06ceecda572a2489345f2299c0f9fa8d
    p177 = 197620364512881247228717050342013327560683201906968909
    assert G2_COFACTOR == 10069 * 5864401 * 1875725156269 * p177 and math.gcd(G2_COFACTOR, R_MOD) == 1
    for p in (10069, 5864401, 1875725156269, p177):
        assert pow(2, p - 1, p) == 1 and pow(3, p - 1, p) == 1          # (Fermat tests; primality proper: tools / DESIGN)
    assert all(G2_COFACTOR % d for d in range(2, 10069))               # 10069 is the smallest prime factor
    assert ref.G2.mul(ref.G2_GEN, R_MOD) is None
    bad = ref.g2_decode(twist_points_outside_subgroup(1, seed=3, small_order=True), False)
    assert ref.G2.on_curve(bad) and ref.G2.mul(bad, R_MOD) is not None
    t = ref.G2.mul(bad, R_MOD)                                         # [r](S + T) = [r]T, of order 10069
    assert ref.G2.mul(t, G2_COFACTOR_SMALL_PRIME) is None
    rng = random.Random(5)
    honest = [ref.G2.mul(ref.G2_GEN, rng.randrange(1, R_MOD)) for _ in range(3)]

    def combo(points, coeffs):
        acc = None
        for p, c in zip(points, coeffs):
            acc = ref.G2.add(acc, ref.G2.mul(p, c))
        return acc
    coeffs = [rng.randrange(1 << 15) for _ in range(4)]
    assert ref.G2.mul(combo(honest, coeffs), R_MOD) is None
    for c in (1, 12345, 10068, 10070, 2 * 10069 + 1):
        assert ref.G2.mul(combo(honest + [bad], coeffs[:3] + [c]), R_MOD) is not None
    assert ref.G2.mul(combo(honest + [bad], coeffs[:3] + [3 * 10069]), R_MOD) is None     # the 1-in-10069 blind spot of ONE sum
