"""The C++ host mirror (include/p2b.hpp): compiles against the C ABI with plain g++ (CPU), fails loudly without a GPU, and
on a GPU reproduces the oracle's bytes for new -> transform -> decompress -> prepare_phase2 -> multiexp -> fft."""
import os
import subprocess

import pytest

from util import R_MOD, be, random_scalars

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
TAU, ALPHA, BETA = 0x1234567 ** 9 % R_MOD, 0x7654321 ** 8 % R_MOD, 0xabcdef1 ** 7 % R_MOD


@pytest.fixture(scope="module")
def exe(tmp_path_factory):
    out = str(tmp_path_factory.mktemp("cpp") / "cpp_mirror_test")
    libdir = os.path.join(ROOT, "phase2_bn254_b200")
    subprocess.check_call(["g++", "-std=c++17", "-O1", "-Wall", "-Wextra", os.path.join(ROOT, "tests", "host", "cpp_mirror_test.cpp"),
                           "-o", out, "-L" + libdir, "-lp2b", "-Wl,-rpath," + libdir])
    return out


def test_cpp_mirror_compiles_and_fails_loudly_without_gpu(exe, tmp_path):
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    (tmp_path / "keys").write_bytes(be(TAU) + be(ALPHA) + be(BETA))
    (tmp_path / "scalars").write_bytes(random_scalars(8, seed=1))
    r = subprocess.run([exe, str(tmp_path), "3"], capture_output=True, text=True)
    assert r.returncode == 3 and "no CUDA device" in r.stderr


@pytest.mark.gpu
def test_cpp_mirror_matches_oracle(exe, tmp_path, oracle):
    size, n = 4, 16
    sc = random_scalars(n, seed=2)
    (tmp_path / "keys").write_bytes(be(TAU) + be(ALPHA) + be(BETA))
    (tmp_path / "scalars").write_bytes(sc)
    r = subprocess.run([exe, str(tmp_path), str(size)], capture_output=True, text=True)
    assert r.returncode == 0, r.stderr
    rd = lambda name: (tmp_path / name).read_bytes()
    ch0 = oracle.pot_generate_initial(size)
    assert rd("challenge")[64:] == ch0[64:]
    exp = oracle.pot_transform(ch0, size, 16, be(TAU), be(ALPHA), be(BETA), threads=4)
    assert rd("response")[64:len(exp)] == exp[64:]
    nxt = oracle.pot_transform(ch0, size, 16, be(TAU), be(ALPHA), be(BETA), out_compressed=False, threads=4)
    assert rd("new_challenge")[64:] == nxt[64:]
    assert rd("msm") == oracle.msm(0, nxt[64:64 + 64 * n], sc, threads=4)
    assert rd("fft") == oracle.fr_fft(sc, threads=2) and rd("fft_roundtrip") == sc
    from util import device_scalars
    rho = device_scalars(bytes((7 * i + 3) & 0xff for i in range(32)), n - 1, 253)
    assert rd("power_pairs") == oracle.msm(0, nxt[64:64 + 64 * (n - 1)], rho, threads=4) + oracle.msm(0, nxt[128:64 + 64 * n], rho, threads=4)
    from phase2_bn254_b200 import lib
    import numpy as np
    c = lib.Context(0)
    assert rd("phase1radix2m") == c.pot_prepare_phase2(np.frombuffer(exp, dtype=np.uint8), size, size, True, True).tobytes()
    c.close()
    # sparse evaluation over tau_g1 (rows: P0 + P1, empty, s0 * P2) and hash_to_g2 of the response's first 32 bytes
    tg1 = nxt[64:]
    assert rd("sparse") == (oracle.sum_points(0, tg1[:128]) + bytes([0x40]) + bytes(63) +
                            oracle.point_mul(0, tg1[128:192], sc[:32]))
    assert rd("hash_to_g2") == lib.hash_to_g2(rd("response")[:32])
