"""GPU parity AT THE SIZES BASELINE.json's configs are quoted on: the CUDA path through the C ABI vs the oracle on identical
inputs, bit for bit.  (The small-size suites cover the edge cases; these cover the code paths that only exist at size:
three-pass FFT plans, multi-group MSM sort/accumulate overlap, multi-chunk host pipelines.)

    config 2  G1 (and G2) Pippenger MSM, 2^20 terms                bellman/src/multiexp.rs:330-475, tests :479-519
    config 3  MPCParameters::contribute, 2^20 constraints          phase2/src/parameters.rs:414-522 (whole output file)
    config 4  Fr FFT / iFFT / coset variants, log n = 17, 18, 20;   bellman/src/domain.rs:154-205, tests :428-457
              one direction at 2^24
    config 1+ BatchedAccumulator::transform, 2^16 powers           powersoftau/src/batched_accumulator.rs:1119-1292

The inputs are valid powers-of-tau vectors produced by the library's own batch_exp (parity-tested separately at small
sizes in test_gpu_batch_mul.py); both sides then consume the same bytes.
"""
import hashlib
import struct

import numpy as np
import pytest

from util import G1_GEN, G2_GEN, R_MOD, be

pytestmark = pytest.mark.gpu
TAU = 0x1d7a3f6c2b9e80415f6a7b8c9d0e1f2031425364758697a8b9cadbecfd0e1f21 % R_MOD
THREADS = 32


def powers(ctx, group, n, start=1):
    """tau^(start+i) * G for i < n (uncompressed wire), on the GPU."""
    gen = np.frombuffer(G2_GEN if group else G1_GEN, dtype=np.uint8)
    return ctx.batch_mul_powers(group, np.tile(gen, n), be(TAU), None, start)


def rand_fr(n, seed):
    rng = np.random.default_rng(seed)
    a = rng.integers(0, 256, size=(n, 32), dtype=np.uint8)
    a[:, 0] &= 0x1f
    return a.reshape(-1)


def full_range_scalars(n, seed):
    """Uniform in [0, r): 254-bit values, the few >= r folded down (exercises the top window)."""
    rng = np.random.default_rng(seed)
    a = rng.integers(0, 256, size=(n, 32), dtype=np.uint8)
    a[:, 0] &= 0x3f
    rb = np.frombuffer(be(R_MOD), dtype=np.uint8).astype(np.int16)
    d = a.astype(np.int16) - rb
    nz = d != 0
    first = nz.argmax(axis=1)
    ge = (d[np.arange(n), first] > 0) | (~nz.any(axis=1))
    a[ge, 0] &= 0x0f
    return a.reshape(-1)


@pytest.mark.parametrize("group", [0, 1])
def test_msm_2p20_matches_oracle(ctx, oracle, group):
    n = 1 << 20
    pts = powers(ctx, group, n)
    sc = full_range_scalars(n, 2020 + group)
    got = ctx.msm(group, pts, sc)
    assert got == oracle.msm(group, pts, sc, threads=THREADS)


@pytest.mark.parametrize("log_n", [17, 18, 20])
def test_fft_three_pass_matches_oracle(ctx, oracle, log_n):
    """log n >= 17 switches fft_plan to three passes with the tmp / tmp2 buffer rotation (csrc/fft.cu)."""
    x = rand_fr(1 << log_n, 1700 + log_n)
    for inv, cos in ((0, 0), (1, 0), (0, 1), (1, 1)):
        assert ctx.fr_fft(x, inv, cos).tobytes() == oracle.fr_fft(x, inv, cos, threads=THREADS), (log_n, inv, cos)


def test_fft_2p24_matches_oracle(ctx, oracle):
    x = rand_fr(1 << 24, 2424)
    y = ctx.fr_fft(x)
    assert hashlib.blake2b(y.tobytes()).digest() == hashlib.blake2b(oracle.fr_fft(x, threads=THREADS)).digest()
    assert np.array_equal(ctx.fr_fft(y, True), x)


def test_contribute_2p20_whole_file(ctx, oracle):
    """Config 3: synthetic MPCParameters with h = 2^20 - 1, l = 2^20; every byte of the output file (2^21 rewritten H / L
    points, delta_g1, delta_g2, the appended public key) and the returned hash against the oracle."""
    nh, nl = (1 << 20) - 1, 1 << 20
    hl = powers(ctx, 0, nh + nl).tobytes()
    g1 = lambda i: hl[64 * i: 64 * i + 64]
    g2 = powers(ctx, 1, 20).tobytes()
    q = lambda i: g2[128 * i: 128 * i + 128]
    params = bytearray()
    params += g1(0) + g1(1) + q(0) + q(1) + g1(2) + q(2)
    params += struct.pack(">I", 2) + g1(3) + g1(4)
    params += struct.pack(">I", nh) + hl[: nh * 64]
    params += struct.pack(">I", nl) + hl[nh * 64: (nh + nl) * 64]
    params += struct.pack(">I", 16) + hl[: 16 * 64]
    params += struct.pack(">I", 16) + hl[64: 17 * 64]
    params += struct.pack(">I", 16) + g2[128 * 3: 128 * 19]
    params += hashlib.blake2b(b"config 3").digest() + struct.pack(">I", 0)
    params = bytes(params)
    delta = be(0x0fedcba987654321fedcba987654321fedcba987654321fedcba987654321 % R_MOD)
    s, r = g1(7), q(19)
    out, h = ctx.phase2_contribute(np.frombuffer(params, dtype=np.uint8), np.frombuffer(delta, dtype=np.uint8),
                                   np.frombuffer(s, dtype=np.uint8), np.frombuffer(r, dtype=np.uint8))
    exp_file, exp_hash = oracle.phase2_contribute(params, delta, s, r, threads=THREADS)
    assert h == exp_hash
    assert hashlib.blake2b(out.tobytes()).digest() == hashlib.blake2b(exp_file).digest()
    assert out.tobytes() == exp_file


def test_transform_2p16_matches_oracle(ctx, oracle):
    """transform on a non-trivial 2^16 challenge (the output of a first GPU contribution, uncompressed), several chunks per
    section, compressed response; whole accumulator region vs the oracle."""
    from phase2_bn254_b200.powersoftau import BatchedAccumulator, CeremonyParams, PrivateKey
    size, batch = 16, 1 << 13
    prm = CeremonyParams(size, batch)
    ch0 = np.frombuffer(oracle.pot_generate_initial(size), dtype=np.uint8)
    ch1 = np.zeros(prm.accumulator_size, dtype=np.uint8)
    key1 = PrivateKey(TAU, 0x2222 * 2**190 % R_MOD, 0x3333 * 2**180 % R_MOD)
    BatchedAccumulator.transform(ch0, ch1, False, False, False, key1, prm, ctx=ctx)
    ch1[:64] = np.frombuffer(hashlib.blake2b(ch0.tobytes()).digest(), dtype=np.uint8)
    key2 = PrivateKey(0x4444 * 2**201 % R_MOD, 0x5555 * 2**170 % R_MOD, TAU)
    rs = np.zeros(prm.contribution_size, dtype=np.uint8)
    BatchedAccumulator.transform(ch1, rs, False, True, True, key2, prm, ctx=ctx)
    exp = oracle.pot_transform(ch1.tobytes(), size, batch, be(key2.tau), be(key2.alpha), be(key2.beta), threads=THREADS)
    end = prm.contribution_size - prm.public_key_size
    assert rs[64:end].tobytes() == exp[64:]
