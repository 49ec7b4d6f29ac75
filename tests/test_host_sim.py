"""CPU: the product's per-point device logic (csrc/fp.cuh, ec.cuh, smul.cuh, codec.cuh, xyzz.cuh are host+device)
compiled for the host by tests/host/host_sim.cpp and compared with the oracle: GLV split, signed fixed windows with
the common-Z table, complete formulas, codecs, XYZZ bucket arithmetic.  No GPU needed."""
import ctypes
import os
import random
import subprocess

import pytest

from util import EDGE_SCALARS, LAMBDA, Q_MOD, R_MOD, be, random_points, random_scalars

HERE = os.path.dirname(os.path.abspath(__file__))


@pytest.fixture(scope="module")
def sim(tmp_path_factory):
    so = str(tmp_path_factory.mktemp("sim") / "libhost_sim.so")
    subprocess.check_call(["g++", "-O2", "-std=c++17", "-x", "c++", "-fPIC", "-shared", "-o", so,
                           os.path.join(HERE, "host", "host_sim.cpp")])
    return ctypes.CDLL(so)


def _mul(sim, group, pt, k, path, in_enc=0, out_enc=0):
    out = (ctypes.c_uint8 * (128 if group else 64))()
    bad = ctypes.c_int(0)
    rc = sim.sim_point_mul(group, pt, k, out, in_enc, out_enc, path, ctypes.byref(bad))
    assert rc == 0
    return bytes(out)[: (128 if group else 64) >> out_enc], bad.value


def test_field_vs_bigint(sim):
    rng = random.Random(21)
    for field, mod in ((0, Q_MOD), (1, R_MOD)):
        vals = [0, 1, mod - 1, 2**255 % mod] + [rng.randrange(mod) for _ in range(300)]
        for a, b in zip(vals, vals[1:]):
            out = (ctypes.c_uint8 * 32)()
            for op, exp in ((0, a * b % mod), (1, (a + b) % mod), (2, (a - b) % mod), (3, pow(a, mod - 2, mod)), (4, -a % mod)):
                sim.sim_field(field, op, be(a), be(b), out)
                assert bytes(out) == be(exp), (field, op, a, b)


def test_glv_split(sim):
    rng = random.Random(22)
    for k in EDGE_SCALARS + [rng.randrange(R_MOD) for _ in range(2000)]:
        k1, k2 = (ctypes.c_uint32 * 5)(), (ctypes.c_uint32 * 5)()
        n1, n2 = ctypes.c_int(0), ctypes.c_int(0)
        sim.sim_glv(be(k), k1, k2, ctypes.byref(n1), ctypes.byref(n2))
        v1 = sum(int(x) << (32 * i) for i, x in enumerate(k1)) * (-1 if n1.value else 1)
        v2 = sum(int(x) << (32 * i) for i, x in enumerate(k2)) * (-1 if n2.value else 1)
        assert (v1 + v2 * LAMBDA - k) % R_MOD == 0
        assert abs(v1) < 2**128 and abs(v2) < 2**128


@pytest.mark.parametrize("group", [0, 1])
def test_scalar_mul_paths_vs_oracle(sim, oracle, group):
    rng = random.Random(23 + group)
    ks = EDGE_SCALARS + [rng.randrange(R_MOD) for _ in range(40 if group == 0 else 12)]
    pts = random_points(oracle, group, len(ks), seed=24)
    size = 128 if group else 64
    exp = oracle.batch_mul(group, pts, b"".join(be(k) for k in ks), threads=8)
    exp_c = oracle.batch_mul(group, pts, b"".join(be(k) for k in ks), 0, 1, threads=8)
    for i, k in enumerate(ks):
        p = pts[i * size:(i + 1) * size]
        # 3 = opt-in endomorphism split on G2 (subgroup points); 4 = uniform-scalar width-5 NAF path (G1, phase-2 shape)
        for path in (0, 1, 2) + ((3,) if group else (4,)):
            got, bad = _mul(sim, group, p, be(k), path)
            if not bad:      # `bad` routes to the complete binary path in the kernel
                assert got == exp[i * size:(i + 1) * size], (group, path, hex(k))
        assert _mul(sim, group, p, be(k), 2)[0] == exp[i * size:(i + 1) * size]
        got_c, bad = _mul(sim, group, p, be(k), 0, 0, 1)
        if not bad:
            assert got_c == exp_c[i * (size // 2):(i + 1) * (size // 2)]


@pytest.mark.parametrize("group", [0, 1])
def test_codecs_vs_oracle(sim, oracle, group):
    size = 128 if group else 64
    pts = random_points(oracle, group, 20, seed=25)
    for i in range(20):
        p = pts[i * size:(i + 1) * size]
        comp = oracle.point_recode(group, p, 0, 1)
        out = (ctypes.c_uint8 * size)()
        assert sim.sim_recode(group, p, out, 0, 1, 1) == 0 and bytes(out)[: size // 2] == comp
        assert sim.sim_recode(group, comp, out, 1, 0, 1) == 0 and bytes(out) == p
    bad = bytearray(pts[:size]); bad[size - 1] ^= 1
    assert sim.sim_recode(group, bytes(bad), out, 0, 0, 1) == oracle.D_NOT_ON_CURVE
    assert sim.sim_recode(group, bytes(bad), out, 0, 0, 0) == 0
    bad = bytearray(pts[:size]); bad[0] |= 0x80
    assert sim.sim_recode(group, bytes(bad), out, 0, 0, 0) == (oracle.D_UNEXPECTED_COMPRESSION if group else oracle.D_UNEXPECTED_INFO)
    bad = bytearray(pts[:size]); bad[0:32] = be(Q_MOD)
    assert sim.sim_recode(group, bytes(bad), out, 0, 0, 0) == oracle.D_COORD


def test_xyzz_arithmetic(sim, oracle):
    n = 12
    pts = bytearray(random_points(oracle, 0, n, seed=26))
    pts[64:128] = pts[0:64]                                   # equal points: doubling branch of the adds
    pts[128:192] = bytes([0x40]) + bytes(63)                  # infinity input
    sc = bytearray(random_scalars(n, seed=27))
    sc[32:64] = sc[0:32]
    out = (ctypes.c_uint8 * 64)()
    assert sim.sim_xyzz_msm(bytes(pts), bytes(sc), n, out) == 0
    assert bytes(out) == oracle.msm(0, bytes(pts), bytes(sc), threads=4)


def test_uniform_scalar_recoding(sim):
    """uniform_digits (csrc/smul.cuh): sum d1[i] 2^i + lambda * sum d2[i] 2^i == k (mod r); digits are odd, |d| <= 15, and
    non-zero digits are at least 5 positions apart (width-5 NAF); the recoding never writes past its 136-digit arrays."""
    rng = random.Random(31)
    ks = EDGE_SCALARS + [rng.randrange(R_MOD) for _ in range(3000)] + [R_MOD - 1 - i for i in range(50)]
    for k in ks:
        buf1, buf2 = (ctypes.c_int8 * 200)(), (ctypes.c_int8 * 200)()
        for i in range(136, 200):
            buf1[i] = buf2[i] = 77                                   # canary behind the 136 digits
        n = sim.sim_uniform_digits(be(k), buf1, buf2)
        assert 0 <= n <= 136
        assert all(buf1[i] == 77 and buf2[i] == 77 for i in range(136, 200))
        for d in (buf1, buf2):
            last = -10
            for i in range(136):
                if d[i]:
                    assert d[i] % 2 != 0 and abs(d[i]) <= 15 and i - last >= 5 and i < n
                    last = i
        v1 = sum(int(buf1[i]) << i for i in range(136))
        v2 = sum(int(buf2[i]) << i for i in range(136))
        assert (v1 + v2 * LAMBDA - k) % R_MOD == 0
