"""Caller buffers as the reference's binaries pass them: memory maps of real files (pageable, file-backed) for the challenge
and the response (powersoftau/src/bin/compute_constrained.rs:83-132), next to page-locked buffers.  Both must give the same
bytes; the pageable ones must have gone through the pinned staging rings of csrc/hostio.cu (p2b_io_stats)."""
import hashlib

import numpy as np
import pytest

from util import G1_GEN, R_MOD, be, random_scalars

pytestmark = pytest.mark.gpu
TAU, ALPHA, BETA = 0x1111 * 2**200 % R_MOD, 0x2222 * 2**190 % R_MOD, 0x3333 * 2**180 % R_MOD


def test_transform_from_and_into_memory_maps(ctx, oracle, tmp_path):
    import torch
    from phase2_bn254_b200.powersoftau import BatchedAccumulator, CeremonyParams, PrivateKey
    size, batch = 12, 1 << 10
    prm = CeremonyParams(size, batch)
    ch = oracle.pot_generate_initial(size)
    (tmp_path / "challenge").write_bytes(ch)
    (tmp_path / "response").write_bytes(bytes(prm.contribution_size))
    cmap = np.memmap(tmp_path / "challenge", dtype=np.uint8, mode="r")
    rmap = np.memmap(tmp_path / "response", dtype=np.uint8, mode="r+")
    key = PrivateKey(TAU, ALPHA, BETA)
    s0 = ctx.io_stats()
    BatchedAccumulator.transform(cmap, rmap, False, True, False, key, prm, ctx=ctx)
    rmap.flush()
    s1 = ctx.io_stats()
    end = prm.contribution_size - prm.public_key_size
    assert s1[0] - s0[0] >= prm.accumulator_size - 64 - 4096 and s1[1] - s0[1] >= end - 64 - 4096   # staged, both directions
    exp = oracle.pot_transform(ch, size, batch, be(TAU), be(ALPHA), be(BETA), threads=8)
    got = (tmp_path / "response").read_bytes()
    assert got[64:end] == exp[64:]
    # the same call on page-locked buffers: direct copies (no staging), same bytes
    pin_c = torch.empty(len(ch), dtype=torch.uint8, pin_memory=True)
    pin_c.numpy()[:] = np.frombuffer(ch, dtype=np.uint8)
    pin_r = torch.zeros(prm.contribution_size, dtype=torch.uint8, pin_memory=True)
    BatchedAccumulator.transform(pin_c.numpy(), pin_r.numpy(), False, True, False, key, prm, ctx=ctx)
    assert ctx.io_stats() == s1
    assert pin_r.numpy()[64:end].tobytes() == exp[64:]


def test_msm_streamed_from_memory_map(ctx, oracle, tmp_path, monkeypatch):
    n = 1 << 16
    pts = ctx.batch_mul_powers(0, np.tile(np.frombuffer(G1_GEN, dtype=np.uint8), n), be(TAU), None, 1)
    sc = random_scalars(n, seed=808)
    (tmp_path / "points").write_bytes(pts.tobytes())
    (tmp_path / "scalars").write_bytes(sc)
    pm = np.memmap(tmp_path / "points", dtype=np.uint8, mode="r")
    sm = np.memmap(tmp_path / "scalars", dtype=np.uint8, mode="r")
    exp = oracle.msm(0, pts, np.frombuffer(sc, dtype=np.uint8), threads=8)
    s0 = ctx.io_stats()
    assert ctx.msm(0, pm, sm) == exp
    assert ctx.io_stats()[0] - s0[0] == 96 * n
    monkeypatch.setenv("P2B_MSM_STREAM_CHUNK", str(1 << 14))      # several streamed chunks, every one staged
    assert ctx.msm(0, pm, sm) == exp


def test_fft_in_place_on_pageable_buffer(ctx, oracle):
    rng = np.random.default_rng(5)
    x = rng.integers(0, 256, size=(1 << 15, 32), dtype=np.uint8)
    x[:, 0] &= 0x1f
    x = x.reshape(-1)
    assert hashlib.blake2b(ctx.fr_fft(x).tobytes()).digest() == hashlib.blake2b(oracle.fr_fft(x, threads=8)).digest()
