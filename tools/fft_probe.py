"""Fr FFT timing (device resident, CUDA events on the library's stream) at a few sizes, with the round-trip check."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import bench
from phase2_bn254_b200 import lib
ctx = lib.Context(0)
dev = torch.device("cuda", 0)
stream = torch.cuda.ExternalStream(ctx.stream_ptr()) if hasattr(ctx, "stream_ptr") else None
for lg in [int(a) for a in sys.argv[1:]] or [16, 20, 22, 24, 26]:
    x = bench.make_scalars(torch, 1 << lg, 0x3237db17, dev)
    y = x.clone()
    ctx.fr_fft_dev(y.data_ptr(), lg, False, False); ctx.fr_fft_dev(y.data_ptr(), lg, True, False); ctx.sync()
    ok = bool(torch.equal(x, y))
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    ctx.profile(True)
    torch.cuda.synchronize()
    reps = 4
    for _ in range(reps):
        ctx.fr_fft_dev(y.data_ptr(), lg, False, False)
        ctx.fr_fft_dev(y.data_ptr(), lg, True, False)
    ctx.sync()
    ms, k = ctx.profile_read(lib.PROF_FFT_PASS)
    ctx.profile(False)
    print("2^%d: %.3f ms per transform in passes (%d pass kernels, %.3f ms each), round trip %s" % (lg, ms / (2 * reps), k, ms / max(1, k), "ok" if ok else "MISMATCH"), flush=True)
