"""CPU: the C oracle (oracle/p2b_oracle.c, the restatement of the reference's algorithms) pinned to the golden
vectors of tests/golden/vectors.json (python big-int arithmetic, tools/make_golden.py) and to the constants the
reference hard-codes."""
import hashlib
import json
import os

import pytest

GOLD = json.load(open(os.path.join(os.path.dirname(__file__), "golden", "vectors.json")))
hx = bytes.fromhex


@pytest.mark.parametrize("group", [0, 1])
def test_batch_mul_golden(oracle, group):
    g = GOLD["g%d_batch_mul" % (group + 1)]
    pts, sc = hx(g["points"]), hx(g["scalars"])
    assert oracle.batch_mul(group, pts, sc).hex() == g["out"]
    assert oracle.batch_mul(group, pts, sc, 0, 1, threads=3).hex() == g["out_compressed"]
    assert oracle.batch_mul(group, hx(g["points_compressed"]), sc, 1, 0).hex() == g["out"]
    b = GOLD["g%d_broadcast" % (group + 1)]
    assert oracle.batch_mul(group, pts, hx(b["scalar"]), threads=2).hex() == b["out"]
    assert oracle.msm(group, pts, sc, threads=4).hex() == GOLD["g%d_msm" % (group + 1)]["out"]


def test_pot_transform_golden(oracle):
    p = GOLD["pot"]
    ch = oracle.pot_generate_initial(p["size"])
    assert hashlib.blake2b(ch).hexdigest() == p["challenge_blake2b"]
    k = [hx(x) for x in p["keys"]]
    for threads in (1, 4):
        r = oracle.pot_transform(ch, p["size"], p["batch"], *k, threads=threads)
        assert r[:64] == hashlib.blake2b(ch).digest()
        assert r[64:].hex() == p["response_body"]
    assert oracle.pot_transform(ch, p["size"], p["batch"], *k, out_compressed=False)[64:].hex() == p["challenge2_body"]
    k2 = [hx(x) for x in p["keys2"]]
    ch2 = bytes(64) + hx(p["challenge2_body"])
    assert oracle.pot_transform(ch2, p["size"], 3, *k2, check_input=True)[64:].hex() == p["response2_body"]
    comp = bytes(64) + hx(p["response_body"])
    assert oracle.pot_transform(comp, p["size"], 8, *k2, in_compressed=True)[64:].hex() == p["response2_body"]


def test_phase2_contribute_golden(oracle):
    p = GOLD["phase2"]
    c = p["c1"]
    out, h = oracle.phase2_contribute(hx(p["params"]), hx(c["delta"]), hx(c["s"]), hx(c["r"]), threads=2)
    assert out.hex() == c["out"] and h.hex() == c["hash"]
    c = p["c2"]
    out2, h2 = oracle.phase2_contribute(out, hx(c["delta"]), hx(c["s"]), hx(c["r"]))
    assert hashlib.blake2b(out2).hexdigest() == c["out_blake2b"] and h2.hex() == c["hash"]


@pytest.mark.parametrize("log_n", [0, 1, 3, 5])
def test_fr_fft_golden(oracle, log_n):
    f = GOLD["fr_fft"][str(log_n)]
    for threads in (1, 4):
        assert oracle.fr_fft(hx(f["in"]), threads=threads).hex() == f["fft"]
        assert oracle.fr_fft(hx(f["in"]), True, threads=threads).hex() == f["ifft"]
        assert oracle.fr_fft(hx(f["in"]), False, True, threads=threads).hex() == f["coset_fft"]
        assert oracle.fr_fft(hx(f["in"]), True, True, threads=threads).hex() == f["icoset_fft"]


def test_reference_constants(oracle):
    """Numbers the reference itself hard-codes (the only stored BN256 values on this path)."""
    c = oracle.constants()
    # R mod q: pairing/src/bn256/fq.rs:39-44
    assert c[0:4] == [0xd35d438dc58f0d9d, 0x0a78eb28f5c70b3d, 0x666ea36f7879462c, 0x0e0a77c19a07df2f]
    q = 21888242871839275222246405745257275088696311157297823662689037894645226208583
    r = 21888242871839275222246405745257275088548364400416034343698204186575808495617
    limbs = lambda v: [(v >> (64 * i)) & (2**64 - 1) for i in range(4)]
    assert c[0:4] == limbs(2**256 % q) and c[4:8] == limbs(2**512 % q)
    assert c[8:12] == limbs(2**256 % r) and c[12:16] == limbs(2**512 % r)
    # blank hash, powersoftau/src/utils.rs:138-140 (first and last bytes as printed there)
    assert GOLD["constants"]["blank_hash"].startswith("786a02f742015903c6c6fd852552d272")
    assert hashlib.blake2b(b"").hexdigest() == GOLD["constants"]["blank_hash"]
    # G1 generator (1, 2): ec.rs:1013-1051; 2G is the well-known alt_bn128 doubling vector
    g = (1).to_bytes(32, "big") + (2).to_bytes(32, "big")
    two_g = oracle.point_mul(0, g, (2).to_bytes(32, "big"))
    assert two_g.hex() == GOLD["constants"]["g1_double"]
    assert int.from_bytes(two_g[:32], "big") == 1368015179489954701390400359078579693043519447331113978918064868415326638035
    assert int.from_bytes(two_g[32:], "big") == 9918110051302171585080402603319702774565515993150576347155970296011118125764
    # Fr: S = 28 (fr.rs:31-34): the root of unity has order exactly 2^28
    w = int(GOLD["constants"]["fr_root_of_unity"], 16)
    assert pow(w, 1 << 28, r) == 1 and pow(w, 1 << 27, r) == r - 1
    # [r]G = infinity, [q - ...]: group order
    assert oracle.point_mul(0, g, (r - 1).to_bytes(32, "big"))[32:] == (q - 2).to_bytes(32, "big")
