"""Shared helpers for the parity tests: seeded synthetic inputs built with the oracle."""
import random

R_MOD = 21888242871839275222246405745257275088548364400416034343698204186575808495617
Q_MOD = 21888242871839275222246405745257275088696311157297823662689037894645226208583
LAMBDA = 0xb3c4d79d41a917585bfc41088d8daaa78b17ea66b99c90dd
G1_GEN = (1).to_bytes(32, "big") + (2).to_bytes(32, "big")
G2_GEN = b"".join(v.to_bytes(32, "big") for v in (
    11559732032986387107991004021392285783925812861821192530917403151452391805634,
    10857046999023057135944570762232829481370756359578518086990519993285655852781,
    4082367875863433681332203403145435568316851327593401208105741076214120093531,
    8495653923123431417604973247489272438418190587263600148770280649306958101930))


def be(x):
    return int(x).to_bytes(32, "big")


EDGE_SCALARS = [0, 1, 2, 3, 15, 16, 17, 31, 32, R_MOD - 1, R_MOD - 2, LAMBDA, LAMBDA - 1, LAMBDA + 1,
                R_MOD - LAMBDA, 1 << 253, (1 << 253) + 1, 1 << 128, (1 << 128) - 1, (1 << 127) + 1]


def random_points(oc, group, n, seed, threads=8):
    """n pseudo-random points [h_i]G (uncompressed wire) via the oracle's batch path."""
    rng = random.Random(seed)
    gen = G2_GEN if group else G1_GEN
    ks = b"".join(be(rng.randrange(1, R_MOD)) for _ in range(n))
    return oc.batch_mul(group, gen * n, ks, threads=threads)


def random_scalars(n, seed):
    rng = random.Random(seed)
    return b"".join(be(rng.randrange(R_MOD)) for _ in range(n))
