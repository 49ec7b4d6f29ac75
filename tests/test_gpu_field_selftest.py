"""The device multipliers against python integers on chosen and random operands: Montgomery product, dedicated squaring, the
fused two-product multiplication (a b + c d) / R that the G1 / G2 addition formulas and the Fq2 product now rest on, and the wide
product + stand-alone reduction -- with the extreme operands (0, 1, p - 1, p - 2, R mod p, 2^k - 1 patterns, all-ones limbs) where
carry-chain mistakes live.  ff_derive's mul_assign / square / add_assign / sub_assign semantics (pairing/src/bn256/fq.rs:4-7,
fr.rs:3-6): canonical residues in, canonical residue out."""
import itertools

import numpy as np
import pytest

from util import Q_MOD, R_MOD

pytestmark = pytest.mark.gpu
RM = 1 << 256


def limbs(vals):
    return np.frombuffer(b"".join(int(v).to_bytes(32, "little") for v in vals), dtype=np.uint8)


def unlimbs(buf):
    b = buf.tobytes()
    return [int.from_bytes(b[i: i + 32], "little") for i in range(0, len(b), 32)]


@pytest.mark.parametrize("field,p", [(0, Q_MOD), (1, R_MOD)])
def test_device_multipliers_on_extreme_and_random_operands(ctx, field, p):
    rinv = pow(RM, -1, p)
    ones = [v for v in ((1 << 32) - 1, (1 << 64) - 1, (1 << 128) - 1, (1 << 224) - 1, (1 << 253) - 1, int("ffffffff" * 7, 16),
                        0x0fffffff_ffffffff_ffffffff_ffffffff_ffffffff_ffffffff_ffffffff_ffffffff) if v < p]
    edge = [0, 1, 2, p - 1, p - 2, p - 3, RM % p, (p - RM % p) % p, (RM * RM) % p, p >> 1, (p >> 1) + 1, 1 << 252, (1 << 253) + 1] + ones
    rng = np.random.default_rng(12345 + field)
    rnd = [int.from_bytes(rng.bytes(32), "little") % p for _ in range(40)]
    vals = edge + rnd
    quads = list(itertools.product(edge, repeat=2))
    a = [x for x, _ in quads]
    b = [y for _, y in quads]
    # c, d: rotate through all values so that every edge pair meets every other one somewhere
    c = [vals[(3 * i + 1) % len(vals)] for i in range(len(a))]
    d = [vals[(5 * i + 2) % len(vals)] for i in range(len(a))]
    # plus a block of purely random quadruples
    m = 20000
    ra = [int.from_bytes(rng.bytes(32), "little") % p for _ in range(m)]
    rb = [int.from_bytes(rng.bytes(32), "little") % p for _ in range(m)]
    rc = [int.from_bytes(rng.bytes(32), "little") % p for _ in range(m)]
    rd = [int.from_bytes(rng.bytes(32), "little") % p for _ in range(m)]
    a, b, c, d = a + ra, b + rb, c + rc, d + rd
    A, B, C, D = limbs(a), limbs(b), limbs(c), limbs(d)
    exp = {0: [x * y * rinv % p for x, y in zip(a, b)],
           1: [x * x * rinv % p for x in a],
           2: [(x * y + z * w) * rinv % p for x, y, z, w in zip(a, b, c, d)],
           3: [x * y * rinv % p for x, y in zip(a, b)],
           4: [(x + y) % p for x, y in zip(a, b)],
           5: [(x - y) % p for x, y in zip(a, b)]}
    for op in range(6):
        if op == 3 and field == 1:
            continue                                   # the stand-alone reduction is only used for Fq
        got = unlimbs(ctx.selftest_field(field, op, A, B, C, D))
        bad = [i for i, (g, e) in enumerate(zip(got, exp[op])) if g != e]
        assert not bad, "op %d: %d mismatches, first at %d: a=%x b=%x c=%x d=%x got=%x exp=%x" % (
            op, len(bad), bad[0], a[bad[0]], b[bad[0]], c[bad[0]], d[bad[0]], got[bad[0]], exp[op][bad[0]])
