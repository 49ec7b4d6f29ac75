"""GPU: the sparse group linear map (p2b_g{1,2}_sparse_mul) and MPCParameters.new on top of it, against the oracle's
restatement of the reference's `eval` loop (phase2/src/parameters.rs:244-300), then the whole ceremony chain the
reference's phase2/test.sh walks: powers of tau -> prepare_phase2 -> new -> contribute -> verify_contribution."""
import hashlib
import random
import struct

import numpy as np
import pytest

from util import G1_GEN, G2_GEN, Q_MOD, R_MOD, be, random_points

pytestmark = pytest.mark.gpu

TAU = 0x1111111111111111111111111111111111111111111111111111111111111111 % R_MOD
ALPHA = 0x2222222222222222222222222222222222222222222222222222222222222222 % R_MOD
BETA = 0x0333333333333333333333333333333333333333333333333333333333333333 % R_MOD


@pytest.mark.parametrize("group", [0, 1])
def test_sparse_mul_vs_oracle(ctx, oracle, group):
    size = 128 if group else 64
    rng = random.Random(40 + group)
    nb = 37
    bases = random_points(oracle, group, nb, seed=41)
    rows = []
    for i in range(60):
        k = rng.choice([0, 1, 1, 2, 3, 5, 9])
        rows.append([(rng.choice([1, R_MOD - 1, 2, rng.randrange(R_MOD)]), rng.randrange(nb)) for _ in range(k)])
    rows.append([(5, 3), (7, 3), (1, 3)])                                  # same base three times
    rows.append([(1, 4), (1, 4)])                                          # P + P: the doubling branch of the mixed add
    rows.append([(9, 5), (R_MOD - 9, 5)])                                  # cancels to infinity
    rows.append([(0, 6)])                                                  # zero coefficient
    rows.append([(rng.randrange(R_MOD), rng.randrange(nb)) for _ in range(700 if group == 0 else 300)])   # > one segment
    offs = np.zeros(len(rows) + 1, dtype=np.uint64)
    offs[1:] = np.cumsum([len(r) for r in rows])
    cols = np.array([c for r in rows for _, c in r], dtype=np.uint32)
    coeffs = b"".join(be(k) for r in rows for k, _ in r)
    got = ctx.sparse_mul(group, bases, offs, cols, coeffs).tobytes()
    exp = oracle.sparse_mul(group, bases, offs, cols, coeffs)
    assert got == exp
    inf = bytes([0x40]) + bytes(size - 1)
    assert got[size * 62: size * 63] == inf and got[size * 63: size * 64] == inf
    assert got[size * 61: size * 62] == oracle.point_mul(group, bases[size * 4: size * 5], be(2))


def test_sparse_mul_edges(ctx, oracle):
    from phase2_bn254_b200 import lib
    # no rows, only empty rows
    assert ctx.sparse_mul(0, G1_GEN, np.zeros(1, dtype=np.uint64), np.zeros(0, dtype=np.uint32), b"").size == 0
    out = ctx.sparse_mul(0, G1_GEN, np.zeros(4, dtype=np.uint64), np.zeros(0, dtype=np.uint32), b"").tobytes()
    assert out == (bytes([0x40]) + bytes(63)) * 3
    # one very long row (three levels of segments): 70,000 x G = [70000]G
    n = 70000
    offs = np.array([0, n, n + 1], dtype=np.uint64)
    out = ctx.sparse_mul(0, G1_GEN, offs, np.zeros(n + 1, dtype=np.uint32), be(1) * (n + 1)).tobytes()
    assert out == oracle.point_mul(0, G1_GEN, be(n)) + G1_GEN
    # errors: column out of range, scalar not canonical, base not on the curve
    with pytest.raises(lib.P2BError) as e:
        ctx.sparse_mul(0, G1_GEN, np.array([0, 1], dtype=np.uint64), np.array([1], dtype=np.uint32), be(1))
    assert e.value.code == lib.EARG
    with pytest.raises(lib.P2BError) as e:
        ctx.sparse_mul(0, G1_GEN, np.array([0, 1], dtype=np.uint64), np.array([0], dtype=np.uint32), be(R_MOD))
    assert e.value.code == lib.EARG
    with pytest.raises(lib.P2BError) as e:
        ctx.sparse_mul(0, G1_GEN[:63] + b"\x03", np.array([0, 1], dtype=np.uint64), np.array([0], dtype=np.uint32), be(1))
    assert e.value.code == lib.EDECODE
    # the context still works afterwards
    assert ctx.sparse_mul(0, G1_GEN, np.array([0, 1], dtype=np.uint64), np.array([0], dtype=np.uint32), be(3)).tobytes() == \
        oracle.point_mul(0, G1_GEN, be(3))


def _ceremony_radix(ctx, size, m):
    """new_constrained -> compute_constrained -> prepare_phase2: the bytes of phase1radix2m{m}."""
    from phase2_bn254_b200.powersoftau import BatchedAccumulator, CeremonyParams, PrivateKey, calculate_hash, prepare_phase2
    params = CeremonyParams(size, 256)
    ch0 = np.zeros(params.accumulator_size, dtype=np.uint8)
    ch0[:64] = np.frombuffer(hashlib.blake2b(b"").digest(), dtype=np.uint8)
    BatchedAccumulator.generate_initial(ch0, False, params)
    rs = np.zeros(params.contribution_size, dtype=np.uint8)
    rs[:64] = np.frombuffer(calculate_hash(ch0), dtype=np.uint8)
    BatchedAccumulator.transform(ch0, rs, False, True, False, PrivateKey(TAU, ALPHA, BETA), params, ctx=ctx)
    end = params.contribution_size - params.public_key_size
    return np.asarray(prepare_phase2(ctx, rs[:end], params, m, input_is_compressed=True)).tobytes()


def _random_circuit(seed, n_pub, n_aux, n_constraints, skip_a_for=None, unconstrained=None):
    def synthesize(cs):
        rng = random.Random(seed)
        one = ("input", 0)
        pub = [cs.alloc_input() for _ in range(n_pub)]
        aux = [cs.alloc() for _ in range(n_aux)]
        allv = [one] + pub + aux
        coeff = lambda: rng.choice([1, 1, -1, 2, rng.randrange(R_MOD)])
        for i in range(n_constraints):
            pick = lambda k: [(rng.choice(allv), coeff()) for _ in range(k)]
            a = [(one, coeff())] + pick(rng.randrange(0, 3))               # ONE in every A row: one very long row
            b = pick(rng.randrange(1, 3))
            c = pick(rng.randrange(0, 2)) + [(aux[i % n_aux], coeff())]    # every aux variable is constrained
            if skip_a_for is not None:
                a = [t for t in a if t[0] != skip_a_for]
            if unconstrained is not None:
                a, b, c = ([t for t in lc if t[0] != unconstrained] for lc in (a, b, c))
            cs.enforce(a, b, c)
    return synthesize


def test_mpc_parameters_new_and_chain(ctx, oracle):
    from phase2_bn254_b200 import lib
    from phase2_bn254_b200.phase2 import (KeypairAssembly, MPCParameters, SynthesisError, params_layout, verify_contribution, _csr)
    n_pub, n_aux, n_con = 2, 40, 300
    radix_cache = {}

    def radix(exp):
        if exp not in radix_cache:
            radix_cache[exp] = _ceremony_radix(ctx, 10, exp)
        return radix_cache[exp]

    circ = _random_circuit(7, n_pub, n_aux, n_con)
    p = MPCParameters.new(circ, False, radix, ctx=ctx)
    lay = params_layout(p.data)
    nvar = 1 + n_pub + n_aux
    assert lay["ic"][1] == 1 + n_pub and lay["l"][1] == n_aux and lay["a"][1] == nvar and lay["b_g2"][1] == nvar
    m = 512
    assert lay["h"][1] == m - 1 and lay["contributions"][1] == 0
    # the oracle's restatement of the eval loop on the same assembly
    cs = KeypairAssembly()
    cs.alloc_input()
    circ(cs)
    for i in range(cs.num_inputs):
        cs.enforce([(("input", i), 1)], [], [])
    assert cs.num_constraints == n_con + 1 + n_pub
    f = radix(9)
    o = 256
    coeffs_g1 = f[o: o + 64 * m]; o += 64 * m
    coeffs_g2 = f[o: o + 128 * m]; o += 128 * m
    alpha_c = f[o: o + 64 * m]; o += 64 * m
    beta_c = f[o: o + 64 * m]; o += 64 * m
    at, bt, ct = cs.at_inputs + cs.at_aux, cs.bt_inputs + cs.bt_aux, cs.ct_inputs + cs.ct_aux
    sect = lambda name: p.section(name).tobytes()
    for name, grp, bases, rows in (("a", 0, coeffs_g1, at), ("b_g1", 0, coeffs_g1, bt), ("b_g2", 1, coeffs_g2, bt)):
        offs, cols, k = _csr(rows)
        assert sect(name) == oracle.sparse_mul(grp, bases, offs, cols, k.tobytes()), name
    ext = b""
    for ra, rb, rc in zip(at, bt, ct):
        parts = []
        for rows, bases in ((ra, beta_c), (rb, alpha_c), (rc, coeffs_g1)):
            offs, cols, k = _csr([rows])
            parts.append(oracle.sparse_mul(0, bases, offs, cols, k.tobytes()))
        ext += oracle.sum_points(0, b"".join(parts))
    assert sect("ic") + sect("l") == ext
    assert sect("h") == f[o: o + 64 * (m - 1)]
    assert sect("alpha_g1") == f[:64] and sect("beta_g1") == f[64:128] and sect("beta_g2") == f[128:256]
    assert sect("gamma_g2") == G2_GEN and sect("delta_g1") == G1_GEN and sect("delta_g2") == G2_GEN
    body_end = lay["cs_hash"][0]
    assert sect("cs_hash") == hashlib.blake2b(p.data[:body_end].tobytes()).digest()
    # the long ONE row went through more than one segment level
    assert len(at[0]) > 256
    # -- contribute on the fresh parameters, then verify (phase2/test.sh)
    p0 = MPCParameters(p.data.copy())
    h1 = p.contribute(rng=lib.ChaChaRng([3] * 8), ctx=ctx)
    assert verify_contribution(p0, p, ctx=ctx, rng=np.random.default_rng(5)) == h1
    # -- filtering points at infinity out of A (a variable that never appears in A)
    skip = ("aux", 5)
    pf = MPCParameters.new(_random_circuit(7, n_pub, n_aux, n_con, skip_a_for=skip), True, radix, ctx=ctx)
    pn = MPCParameters.new(_random_circuit(7, n_pub, n_aux, n_con, skip_a_for=skip), False, radix, ctx=ctx)
    assert params_layout(pf.data)["a"][1] == nvar - 1 and params_layout(pn.data)["a"][1] == nvar
    a_n = pn.section("a").reshape(-1, 64)
    assert a_n[1 + n_pub + 5, 0] == 0x40
    assert pf.section("a").tobytes() == np.delete(a_n, 1 + n_pub + 5, axis=0).tobytes()
    # -- an aux variable in no constraint at all: UnconstrainedVariable
    with pytest.raises(SynthesisError):
        MPCParameters.new(_random_circuit(7, n_pub, n_aux, n_con, unconstrained=("aux", 9)), False, radix, ctx=ctx)
