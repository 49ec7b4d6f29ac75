"""GPU: the device-pointer entry points (`*_dev`: inputs and outputs resident in HBM, asynchronous on p2b_stream until
p2b_sync) give the same bytes as the host-buffer entry points, and the profiling hooks count their kernels."""
import numpy as np
import pytest
import torch

from util import R_MOD, be, random_points, random_scalars

pytestmark = pytest.mark.gpu


def dev(a):
    return torch.from_numpy(np.frombuffer(bytes(a), dtype=np.uint8).copy()).cuda()


@pytest.mark.parametrize("group", [0, 1])
def test_batch_mul_dev_matches_host(ctx, oracle, group):
    from phase2_bn254_b200 import lib
    n, size = 777, 128 if group else 64
    pts, sc = random_points(oracle, group, n, seed=801), random_scalars(n, seed=802)
    d_in = dev(pts)
    d_out = torch.empty(n * size // 2, dtype=torch.uint8, device="cuda")
    ctx.batch_mul_dev(group, d_in.data_ptr(), d_out.data_ptr(), n, np.frombuffer(sc, dtype=np.uint8), 0, 1)
    ctx.sync()
    assert d_out.cpu().numpy().tobytes() == ctx.batch_mul(group, pts, sc, 0, 1).tobytes()
    k = be(0x77665544332211 ** 3 % R_MOD)
    d_out2 = torch.empty(n * size, dtype=torch.uint8, device="cuda")
    ctx.batch_mul_dev(group, d_in.data_ptr(), d_out2.data_ptr(), n, np.frombuffer(k, dtype=np.uint8))
    ctx.sync()
    assert d_out2.cpu().numpy().tobytes() == oracle.batch_mul(group, pts, k, threads=8)
    tau = be(0x1234567 ** 9 % R_MOD)
    ctx.batch_mul_powers_dev(group, d_in.data_ptr(), d_out2.data_ptr(), n, np.frombuffer(tau, dtype=np.uint8), None, 3)
    ctx.sync()
    assert d_out2.cpu().numpy().tobytes() == oracle.batch_mul_powers(group, pts, tau, None, 3, threads=8)
    # errors surface at sync()
    bad = bytearray(pts); bad[size * 5] |= 0x80
    ctx.batch_mul_dev(group, dev(bad).data_ptr(), d_out2.data_ptr(), n, np.frombuffer(k, dtype=np.uint8))
    with pytest.raises(lib.P2BError) as e:
        ctx.sync()
    assert e.value.code == lib.EDECODE and e.value.index == 5
    ctx.sync()                                                    # the error word is reset


def test_msm_and_fft_dev_match_host(ctx, oracle):
    from phase2_bn254_b200 import lib
    n = 3000
    pts, sc = random_points(oracle, 0, n, seed=803), random_scalars(n, seed=804)
    d_p, d_s = dev(pts), dev(sc)
    ctx.profile(True)
    assert ctx.msm_dev(0, d_p.data_ptr(), d_s.data_ptr(), n) == oracle.msm(0, pts, sc, threads=8)
    ms, kernels = ctx.profile_read(lib.PROF_MSM_ACCUMULATE)
    assert kernels >= 1 and ms > 0
    ctx.profile(False)
    x = random_scalars(1 << 12, seed=805)
    d_x = dev(x)
    ctx.fr_fft_dev(d_x.data_ptr(), 12, False, True)
    ctx.sync()
    assert d_x.cpu().numpy().tobytes() == oracle.fr_fft(x, False, True, threads=8)
    ctx.fr_fft_dev(d_x.data_ptr(), 12, True, True)
    ctx.sync()
    assert d_x.cpu().numpy().tobytes() == x
    before = ctx.launch_count
    ctx.fr_fft_dev(d_x.data_ptr(), 12, False, False)
    ctx.sync()
    assert ctx.launch_count > before


@pytest.mark.parametrize("group", [0, 1])
def test_msm_dev_tiled(ctx, oracle, group, monkeypatch):
    """Device-resident inputs above the tile size are accumulated tile by tile into the same buckets (L2 locality of the
    gathers); the hook lowers the tile size so that the path runs at test sizes, including a ragged last tile."""
    n = 2500
    pts, sc = random_points(oracle, group, n, seed=811 + group), random_scalars(n, seed=812)
    d_p, d_s = dev(pts), dev(sc)
    exp = oracle.msm(group, pts, sc, threads=8)
    for tile in ("700", "1024", "2499", "0"):
        monkeypatch.setenv("P2B_MSM_TILE", tile)
        assert ctx.msm_dev(group, d_p.data_ptr(), d_s.data_ptr(), n) == exp, tile
