"""ctypes binding of libp2b.so (include/p2b.h) -- the only way into the compute core from Python.

There is no CPU fallback: if the shared library is missing or no CUDA device is present, the
constructors raise.  Buffers are numpy uint8 arrays (host) or raw device pointers (ints).
"""
import ctypes
import os

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("P2B_LIB") or os.path.join(_HERE, "libp2b.so")   # P2B_LIB: tuning builds of the same library

OK, EARG, EDECODE, EINFINITY_IN, EINFINITY_OUT, ECUDA = range(6)
DEC_NOT_ON_CURVE, DEC_COORDINATE, DEC_UNEXPECTED_INFORMATION, DEC_UNEXPECTED_COMPRESSION_MODE = 1, 2, 3, 4
ENC_UNCOMPRESSED, ENC_COMPRESSED, ENC_RAW_MONT_LE = 0, 1, 2
CHECK_INPUT, REJECT_INFINITY, G2_SUBGROUP, G2_EXACT = 1, 2, 4, 8
G1, G2 = 0, 1

# every symbol include/p2b.h declares (tests/test_abi.py checks the header and this list agree)
SYMBOLS = [
    "p2b_init", "p2b_destroy", "p2b_last_error", "p2b_error_detail", "p2b_stream", "p2b_launch_count",
    "p2b_version", "p2b_g1_batch_mul", "p2b_g2_batch_mul", "p2b_g1_batch_mul_powers",
    "p2b_g2_batch_mul_powers", "p2b_g1_batch_mul_dev", "p2b_g2_batch_mul_dev",
    "p2b_g1_batch_mul_powers_dev", "p2b_g2_batch_mul_powers_dev", "p2b_sync",
    "p2b_pot_accumulator_size", "p2b_pot_transform", "p2b_phase2_transcript", "p2b_phase2_contribute",
    "p2b_g1_msm", "p2b_g2_msm", "p2b_g1_msm_dev", "p2b_g2_msm_dev", "p2b_g1_sum_points", "p2b_g2_sum_points",
    "p2b_fr_fft", "p2b_fr_fft_dev", "p2b_profile_enable", "p2b_profile_read",
    "p2b_pot_decompress", "p2b_g1_recode", "p2b_g2_recode",
    "p2b_g1_group_fft", "p2b_g2_group_fft", "p2b_pot_radix_file_size", "p2b_pot_prepare_phase2",
    "p2b_g1_sparse_mul", "p2b_g2_sparse_mul",
    "p2b_pairing_check", "p2b_same_ratio", "p2b_hash_to_g2", "p2b_rng_seed", "p2b_rng_u32", "p2b_rng_fr", "p2b_rng_g1",
    "p2b_rng_g2", "p2b_host_g1_mul", "p2b_host_g2_mul", "p2b_pairing_constants", "p2b_io_stats",
    "p2b_g1_msm_pair", "p2b_g2_msm_pair", "p2b_g1_power_pairs", "p2b_g2_power_pairs", "p2b_random_scalars", "p2b_phase2_contribute_sharded",
    "p2b_g1_group_fft_scaled", "p2b_g2_group_fft_scaled", "p2b_g1_gfft_stage", "p2b_g2_gfft_stage", "p2b_fr_root_of_unity", "p2b_selftest_field",
    "p2b_g2_probe_stats",
]
PROF_BATCH_MUL, PROF_NORMALIZE, PROF_MSM_SORT, PROF_MSM_ACCUMULATE, PROF_MSM_REDUCE, PROF_FFT_PASS = range(6)


class P2BError(RuntimeError):
    """A non-zero return code from the C ABI (code, decode sub-code, first failing element)."""

    def __init__(self, code, message, index=0, sub=0):
        super().__init__("p2b error %d: %s" % (code, message))
        self.code, self.index, self.sub = code, index, sub


_lib = None


def load():
    """dlopen libp2b.so; raises ImportError when it has not been built (no fallback)."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise ImportError("libp2b.so is not built: run `python -c 'import __graft_entry__ as g; g.build()'` "
                          "or `make -C phase2_bn254_b200/csrc`")
    lib = ctypes.CDLL(LIB_PATH)
    vp, u8p, sz, u64, i32, u32 = (ctypes.c_void_p, ctypes.c_void_p, ctypes.c_size_t, ctypes.c_uint64, ctypes.c_int,
                                  ctypes.c_uint32)
    lib.p2b_init.argtypes = [i32, ctypes.POINTER(vp)]
    lib.p2b_destroy.argtypes = [vp]
    lib.p2b_destroy.restype = None
    lib.p2b_last_error.argtypes = [vp]
    lib.p2b_last_error.restype = ctypes.c_char_p
    lib.p2b_error_detail.argtypes = [vp, ctypes.POINTER(u64), ctypes.POINTER(i32)]
    lib.p2b_error_detail.restype = None
    lib.p2b_stream.argtypes = [vp]
    lib.p2b_stream.restype = vp
    lib.p2b_launch_count.argtypes = [vp]
    lib.p2b_launch_count.restype = u64
    lib.p2b_version.restype = ctypes.c_char_p
    lib.p2b_g2_probe_stats.argtypes = [vp, ctypes.POINTER(u64), ctypes.POINTER(i32)]
    for g in ("g1", "g2"):
        for suffix in ("", "_dev"):
            getattr(lib, "p2b_%s_batch_mul%s" % (g, suffix)).argtypes = [vp, u8p, u8p, sz, u8p, sz, i32, i32, i32]
            getattr(lib, "p2b_%s_batch_mul_powers%s" % (g, suffix)).argtypes = [vp, u8p, u8p, sz, u8p, u8p, u64, i32,
                                                                               i32, i32]
        if hasattr(lib, "p2b_%s_msm" % g):
            getattr(lib, "p2b_%s_msm" % g).argtypes = [vp, u8p, u8p, sz, u8p]
            getattr(lib, "p2b_%s_msm_dev" % g).argtypes = [vp, vp, vp, sz, u8p]
            getattr(lib, "p2b_%s_sum_points" % g).argtypes = [vp, u8p, sz, u8p]
        getattr(lib, "p2b_%s_msm_pair" % g).argtypes = [vp, u8p, u8p, u8p, sz, u8p, u32, i32, i32, u8p, u8p]
        getattr(lib, "p2b_%s_power_pairs" % g).argtypes = [vp, u8p, sz, u8p, u8p, u32, i32, i32, u8p, u8p]
    lib.p2b_random_scalars.argtypes = [vp, u8p, u64, sz, u32, u8p]
    lib.p2b_sync.argtypes = [vp]
    lib.p2b_pot_accumulator_size.argtypes = [u32, i32]
    lib.p2b_pot_accumulator_size.restype = u64
    lib.p2b_pot_transform.argtypes = [vp, u8p, u64, u8p, u64, u32, u32, i32, i32, i32, u8p, u8p, u8p, u32, u32]
    lib.p2b_phase2_transcript.argtypes = [vp, u8p, u64, u8p, u8p, u8p]
    lib.p2b_phase2_contribute.argtypes = [vp, u8p, u64, u8p, u64, u8p, u8p, u8p, u8p]
    lib.p2b_phase2_contribute_sharded.argtypes = [vp, u8p, u64, u8p, u64, u8p, u8p, u8p, u8p, u32, u32]
    if hasattr(lib, "p2b_fr_fft"):
        lib.p2b_fr_fft.argtypes = [vp, u8p, u32, i32, i32]
        lib.p2b_fr_fft_dev.argtypes = [vp, vp, u32, i32, i32]
    lib.p2b_pot_decompress.argtypes = [vp, u8p, u64, u8p, u64, u32, i32, u32, u32]
    lib.p2b_g1_recode.argtypes = [vp, u8p, u8p, sz, i32, i32, i32]
    lib.p2b_g2_recode.argtypes = [vp, u8p, u8p, sz, i32, i32, i32]
    lib.p2b_g1_group_fft.argtypes = [vp, u8p, u8p, u32, i32, i32, i32, i32]
    lib.p2b_g2_group_fft.argtypes = [vp, u8p, u8p, u32, i32, i32, i32, i32]
    for g in ("g1", "g2"):
        getattr(lib, "p2b_%s_group_fft_scaled" % g).argtypes = [vp, u8p, u8p, u32, i32, i32, i32, i32, u32]
        getattr(lib, "p2b_%s_gfft_stage" % g).argtypes = [vp, u8p, u8p, u32, u8p, u64, i32, i32, i32, u8p, u8p]
    lib.p2b_fr_root_of_unity.argtypes = [u32, i32, u8p]
    lib.p2b_selftest_field.argtypes = [vp, i32, i32, u8p, u8p, u8p, u8p, sz, u8p]
    lib.p2b_pot_radix_file_size.argtypes = [u32]
    lib.p2b_pot_radix_file_size.restype = u64
    lib.p2b_pot_prepare_phase2.argtypes = [vp, u8p, u64, u32, i32, i32, u32, u8p, u64, i32]
    lib.p2b_g1_sparse_mul.argtypes = [vp, u8p, sz, vp, vp, u8p, sz, u8p]
    lib.p2b_g2_sparse_mul.argtypes = [vp, u8p, sz, vp, vp, u8p, sz, u8p]
    lib.p2b_pairing_check.argtypes = [u8p, u8p, sz, ctypes.POINTER(i32)]
    lib.p2b_same_ratio.argtypes = [u8p, u8p, u8p, u8p, ctypes.POINTER(i32)]
    lib.p2b_hash_to_g2.argtypes = [u8p, u8p]
    lib.p2b_rng_seed.argtypes = [u8p, ctypes.POINTER(u32)]
    lib.p2b_rng_u32.argtypes = [u8p, ctypes.POINTER(u32)]
    for name in ("p2b_rng_fr", "p2b_rng_g1", "p2b_rng_g2", "p2b_pairing_constants"):
        getattr(lib, name).argtypes = [u8p] * (1 if name == "p2b_pairing_constants" else 2)
    lib.p2b_host_g1_mul.argtypes = [u8p, u8p, u8p]
    lib.p2b_host_g2_mul.argtypes = [u8p, u8p, u8p]
    lib.p2b_io_stats.argtypes = [vp, ctypes.POINTER(u64), ctypes.POINTER(u64)]
    lib.p2b_io_stats.restype = None
    lib.p2b_profile_enable.argtypes = [vp, i32]
    lib.p2b_profile_read.argtypes = [vp, i32, ctypes.POINTER(ctypes.c_double), ctypes.POINTER(u64)]
    _lib = lib
    return lib


def enc_size(group, enc):
    full = 128 if group == G2 else 64
    return full // 2 if enc == ENC_COMPRESSED else full


def _host(buf):
    """numpy uint8 view of a bytes-like / array (no copy when already contiguous)."""
    if isinstance(buf, np.ndarray):
        a = buf if buf.dtype == np.uint8 else buf.view(np.uint8)
        return np.ascontiguousarray(a).reshape(-1)
    return np.frombuffer(buf, dtype=np.uint8)


def _ptr(a):
    return ctypes.c_void_p(a.ctypes.data) if isinstance(a, np.ndarray) else ctypes.c_void_p(int(a))


# ---- verifier host side (CPU code inside libp2b.so, as in the reference: pairings, hash_to_g2, key-generation RNG) ----
def _hostcall(name, *args):
    rc = getattr(load(), name)(*args)
    if rc:
        raise P2BError(rc, {EARG: "bad argument", EDECODE: "point does not decode"}.get(rc, "error") + " in " + name)


def _fixed(buf, size, what):
    a = _host(buf)
    if a.size != size:
        raise ValueError("%s must be %d bytes, got %d" % (what, size, a.size))
    return a


def pairing_check(g1_points, g2_points):
    """prod_i e(P_i, Q_i) == 1 over uncompressed points (pairs with a point at infinity contribute 1)."""
    a, b = _host(g1_points), _host(g2_points)
    if a.size % 64 or b.size != 2 * a.size:
        raise ValueError("need n G1 points (64 B) and n G2 points (128 B)")
    r = ctypes.c_int(0)
    _hostcall("p2b_pairing_check", _ptr(a), _ptr(b), a.size // 64, ctypes.byref(r))
    return bool(r.value)


def same_ratio(g1, g2):
    """same_ratio((a, b), (c, d)): e(a, d) == e(b, c) (powersoftau/src/utils.rs:151-159, phase2/src/utils.rs:48-57)."""
    a, b = _fixed(g1[0], 64, "g1.0"), _fixed(g1[1], 64, "g1.1")
    c, d = _fixed(g2[0], 128, "g2.0"), _fixed(g2[1], 128, "g2.1")
    r = ctypes.c_int(0)
    _hostcall("p2b_same_ratio", _ptr(a), _ptr(b), _ptr(c), _ptr(d), ctypes.byref(r))
    return bool(r.value)


def hash_to_g2(digest):
    """hash_to_g2(digest) (powersoftau/src/utils.rs:31-45): 128-byte uncompressed G2 point from the first 32 bytes."""
    d = _host(digest)
    if d.size < 32:
        raise ValueError("digest must be at least 32 bytes")
    d = np.ascontiguousarray(d[:32])
    out = np.empty(128, dtype=np.uint8)
    _hostcall("p2b_hash_to_g2", _ptr(d), _ptr(out))
    return out.tobytes()


def root_of_unity(log_d, inverse=False):
    """omega of the 2^log_d evaluation domain (bellman/src/domain.rs:52-99) as an int; inverse: omega^-1."""
    out = np.empty(32, dtype=np.uint8)
    _hostcall("p2b_fr_root_of_unity", int(log_d), int(bool(inverse)), _ptr(out))
    return int.from_bytes(out.tobytes(), "big")


def host_mul(group, point, scalar_be32):
    """One scalar multiplication on the host (CurveAffine::mul of the key generation, keypair.rs:64-84)."""
    size = 128 if group == G2 else 64
    p, k = _fixed(point, size, "point"), _fixed(scalar_be32, 32, "scalar")
    out = np.empty(size, dtype=np.uint8)
    _hostcall("p2b_host_g2_mul" if group == G2 else "p2b_host_g1_mul", _ptr(p), _ptr(k), _ptr(out))
    return out.tobytes()


class ChaChaRng:
    """rand 0.4.6 `ChaChaRng::from_seed(&[u32; 8])` with the reference's samplers (restated; the crate is not vendored)."""

    def __init__(self, seed_words):
        words = [int(w) & 0xffffffff for w in seed_words]
        if len(words) != 8:
            raise ValueError("seed is 8 u32 words")
        self._state = np.zeros(136, dtype=np.uint8)
        _hostcall("p2b_rng_seed", _ptr(self._state), (ctypes.c_uint32 * 8)(*words))

    @classmethod
    def from_digest(cls, digest):
        """Seed from the first 32 bytes as 8 big-endian words (hash_to_g2, beacon_constrained.rs:81-93)."""
        d = bytes(digest)[:32]
        return cls([int.from_bytes(d[4 * i: 4 * i + 4], "big") for i in range(8)])

    def next_u32(self):
        v = ctypes.c_uint32(0)
        _hostcall("p2b_rng_u32", _ptr(self._state), ctypes.byref(v))
        return v.value

    def _gen(self, name, size):
        out = np.empty(size, dtype=np.uint8)
        _hostcall(name, _ptr(self._state), _ptr(out))
        return out.tobytes()

    def gen_fr(self):
        """Fr::rand as an int in [0, r)."""
        return int.from_bytes(self._gen("p2b_rng_fr", 32), "big")

    def gen_g1(self):
        return self._gen("p2b_rng_g1", 64)

    def gen_g2(self):
        return self._gen("p2b_rng_g2", 128)


class Context:
    """One compute context on one GPU (p2b_init / p2b_destroy)."""

    def __init__(self, device=0):
        self.lib = load()
        h = ctypes.c_void_p()
        rc = self.lib.p2b_init(int(device), ctypes.byref(h))
        if rc != OK:
            raise P2BError(rc, "p2b_init failed on device %d (no CUDA device / driver?)" % device)
        self.h = h
        self.device = device

    def close(self):
        if getattr(self, "h", None):
            self.lib.p2b_destroy(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    # -- helpers
    def _check(self, rc):
        if rc != OK:
            idx, sub = ctypes.c_uint64(0), ctypes.c_int(0)
            self.lib.p2b_error_detail(self.h, ctypes.byref(idx), ctypes.byref(sub))
            raise P2BError(rc, self.lib.p2b_last_error(self.h).decode(), idx.value, sub.value)

    @property
    def stream(self):
        return self.lib.p2b_stream(self.h)

    @property
    def launch_count(self):
        return self.lib.p2b_launch_count(self.h)

    def g2_probe_stats(self):
        """(batches probed so far, verdict of the last probe: 0 = subgroup proven / split path, 1 = exact path, -1 = none)"""
        n, v = ctypes.c_uint64(0), ctypes.c_int(0)
        self._check(self.lib.p2b_g2_probe_stats(self.h, ctypes.byref(n), ctypes.byref(v)))
        return n.value, v.value

    def sync(self):
        self._check(self.lib.p2b_sync(self.h))

    def io_stats(self):
        """(H2D bytes, D2H bytes) of caller buffers that were pageable and went through the pinned staging rings."""
        a, b = ctypes.c_uint64(0), ctypes.c_uint64(0)
        self.lib.p2b_io_stats(self.h, ctypes.byref(a), ctypes.byref(b))
        return a.value, b.value

    def profile(self, on=True):
        """Bracket the dominant kernels with CUDA events on the ctx stream (and reset the counters)."""
        self._check(self.lib.p2b_profile_enable(self.h, int(on)))

    def profile_read(self, slot):
        """(total device ms, kernel launches) recorded in `slot` since profile(True)."""
        ms, k = ctypes.c_double(0), ctypes.c_uint64(0)
        self._check(self.lib.p2b_profile_read(self.h, slot, ctypes.byref(ms), ctypes.byref(k)))
        return ms.value, k.value

    # -- level 1 (host buffers)
    def batch_mul(self, group, points, scalars, in_enc=ENC_UNCOMPRESSED, out_enc=ENC_UNCOMPRESSED, flags=0, out=None):
        pts, sc = _host(points), _host(scalars)
        n = pts.size // enc_size(group, in_enc)
        if out is None:
            out = np.empty(max(1, n * enc_size(group, out_enc)), dtype=np.uint8)
        fn = self.lib.p2b_g2_batch_mul if group == G2 else self.lib.p2b_g1_batch_mul
        self._check(fn(self.h, _ptr(pts), _ptr(out), n, _ptr(sc), sc.size // 32, in_enc, out_enc, flags))
        return out[: n * enc_size(group, out_enc)]

    def batch_mul_powers(self, group, points, tau, coeff=None, start=0, in_enc=ENC_UNCOMPRESSED,
                         out_enc=ENC_UNCOMPRESSED, flags=REJECT_INFINITY, out=None):
        pts = _host(points)
        n = pts.size // enc_size(group, in_enc)
        if out is None:
            out = np.empty(max(1, n * enc_size(group, out_enc)), dtype=np.uint8)
        fn = self.lib.p2b_g2_batch_mul_powers if group == G2 else self.lib.p2b_g1_batch_mul_powers
        t = _host(tau)
        cf = _host(coeff) if coeff is not None else None
        self._check(fn(self.h, _ptr(pts), _ptr(out), n, _ptr(t), _ptr(cf) if cf is not None else None, start, in_enc,
                       out_enc, flags))
        return out[: n * enc_size(group, out_enc)]

    def recode(self, group, points, in_enc, out_enc, flags=0, out=None):
        """Bulk codec: decompress / compress / checked deserialisation without a scalar multiplication."""
        pts = _host(points)
        n = pts.size // enc_size(group, in_enc)
        if out is None:
            out = np.empty(max(1, n * enc_size(group, out_enc)), dtype=np.uint8)
        fn = self.lib.p2b_g2_recode if group == G2 else self.lib.p2b_g1_recode
        self._check(fn(self.h, _ptr(pts), _ptr(out), n, in_enc, out_enc, flags))
        return out[: n * enc_size(group, out_enc)]

    def validate(self, group, points, in_enc=ENC_UNCOMPRESSED, flags=CHECK_INPUT):
        """Decode + check every point on the device without copying anything back (raises P2BError like recode)."""
        pts = _host(points)
        n = pts.size // enc_size(group, in_enc)
        fn = self.lib.p2b_g2_recode if group == G2 else self.lib.p2b_g1_recode
        self._check(fn(self.h, _ptr(pts), None, n, in_enc, ENC_UNCOMPRESSED, flags))

    def pot_decompress(self, response, challenge, size_log2, check_input=False, shard_index=0, shard_count=1):
        rs, ch = _host(response), challenge
        assert isinstance(ch, np.ndarray) and ch.dtype == np.uint8 and ch.flags.writeable
        self._check(self.lib.p2b_pot_decompress(self.h, _ptr(rs), rs.size, _ptr(ch), ch.size, size_log2, int(check_input),
                                                shard_index, shard_count))

    # -- level 1 (device pointers; asynchronous until sync())
    def batch_mul_dev(self, group, d_in, d_out, n, scalars, in_enc=ENC_UNCOMPRESSED, out_enc=ENC_UNCOMPRESSED, flags=0):
        sc = _host(scalars)
        fn = self.lib.p2b_g2_batch_mul_dev if group == G2 else self.lib.p2b_g1_batch_mul_dev
        self._check(fn(self.h, _ptr(d_in), _ptr(d_out), n, _ptr(sc), sc.size // 32, in_enc, out_enc, flags))

    def batch_mul_powers_dev(self, group, d_in, d_out, n, tau, coeff=None, start=0, in_enc=ENC_UNCOMPRESSED,
                             out_enc=ENC_UNCOMPRESSED, flags=REJECT_INFINITY):
        fn = self.lib.p2b_g2_batch_mul_powers_dev if group == G2 else self.lib.p2b_g1_batch_mul_powers_dev
        t = _host(tau)
        cf = _host(coeff) if coeff is not None else None
        self._check(fn(self.h, _ptr(d_in), _ptr(d_out), n, _ptr(t), _ptr(cf) if cf is not None else None, start, in_enc,
                       out_enc, flags))

    # -- level 2
    def pot_accumulator_size(self, size_log2, compressed):
        return self.lib.p2b_pot_accumulator_size(size_log2, int(compressed))

    def pot_transform(self, challenge, response, size_log2, batch_size, tau, alpha, beta, in_compressed=False,
                      out_compressed=True, check_input=False, shard_index=0, shard_count=1):
        ch, rs = _host(challenge), response
        assert isinstance(rs, np.ndarray) and rs.dtype == np.uint8 and rs.flags.writeable
        self._check(self.lib.p2b_pot_transform(self.h, _ptr(ch), ch.size, _ptr(rs), rs.size, size_log2, batch_size,
                                               int(in_compressed), int(out_compressed), int(check_input),
                                               _ptr(_host(tau)), _ptr(_host(alpha)), _ptr(_host(beta)), shard_index,
                                               shard_count))

    def phase2_transcript(self, params, delta, s_g1):
        p = _host(params)
        out = np.empty(64, dtype=np.uint8)
        self._check(self.lib.p2b_phase2_transcript(self.h, _ptr(p), p.size, _ptr(_host(delta)), _ptr(_host(s_g1)),
                                                   _ptr(out)))
        return out.tobytes()

    def phase2_contribute(self, params, delta, s_g1, r_g2, out=None, shard_index=0, shard_count=1):
        p = _host(params)
        if out is None:
            out = np.empty(p.size + 384, dtype=np.uint8)
        h = np.empty(64, dtype=np.uint8)
        self._check(self.lib.p2b_phase2_contribute_sharded(self.h, _ptr(p), p.size, _ptr(out), out.size, _ptr(_host(delta)),
                                                           _ptr(_host(s_g1)), _ptr(_host(r_g2)), _ptr(h), shard_index, shard_count))
        return out[: p.size + 384], h.tobytes()

    # -- MSM
    def msm(self, group, points, scalars):
        pts, sc = _host(points), _host(scalars)
        n = sc.size // 32
        if sc.size % 32 or pts.size != n * enc_size(group, ENC_UNCOMPRESSED):
            raise ValueError("msm: %d scalar bytes need %d point bytes, got %d" % (sc.size, n * enc_size(group, ENC_UNCOMPRESSED),
                                                                                pts.size))
        out = np.empty(enc_size(group, ENC_UNCOMPRESSED), dtype=np.uint8)
        fn = self.lib.p2b_g2_msm if group == G2 else self.lib.p2b_g1_msm
        self._check(fn(self.h, _ptr(pts), _ptr(sc), n, _ptr(out)))
        return out.tobytes()

    def msm_pair(self, group, points_a, points_b, scalars=None, seed=None, scalar_bits=0, in_enc=ENC_UNCOMPRESSED, flags=0):
        """(sum k_i A_i, sum k_i B_i) in one pass (merge_pairs).  scalars=None: coefficients generated on the device from the
        32-byte `seed` (ChaCha20 keystream), scalar_bits bits each."""
        pa, pb = _host(points_a), _host(points_b)
        size = enc_size(group, in_enc)
        n = pa.size // size
        if pa.size % size or pb.size != pa.size:
            raise ValueError("msm_pair: the two point arrays must hold the same number of points")
        sc = _host(scalars) if scalars is not None else None
        if sc is not None and sc.size != 32 * n:
            raise ValueError("msm_pair: need %d scalar bytes, got %d" % (32 * n, sc.size))
        sd = _fixed(seed, 32, "seed") if seed is not None else None
        full = enc_size(group, ENC_UNCOMPRESSED)
        oa, ob = np.empty(full, dtype=np.uint8), np.empty(full, dtype=np.uint8)
        fn = self.lib.p2b_g2_msm_pair if group == G2 else self.lib.p2b_g1_msm_pair
        self._check(fn(self.h, _ptr(pa), _ptr(pb), _ptr(sc) if sc is not None else None, n, _ptr(sd) if sd is not None else None,
                       scalar_bits, in_enc, flags, _ptr(oa), _ptr(ob)))
        return oa.tobytes(), ob.tobytes()

    def power_pairs(self, group, points, scalars=None, seed=None, scalar_bits=0, in_enc=ENC_UNCOMPRESSED, flags=0):
        """merge_pairs(v[..n-1], v[1..]) over n points in one pass (power_pairs, powersoftau/src/utils.rs:133-135)."""
        p = _host(points)
        size = enc_size(group, in_enc)
        n = p.size // size
        if p.size % size or n < 1:
            raise ValueError("power_pairs: need at least one point")
        sc = _host(scalars) if scalars is not None else None
        if sc is not None and sc.size != 32 * (n - 1):
            raise ValueError("power_pairs: need %d scalar bytes, got %d" % (32 * (n - 1), sc.size))
        sd = _fixed(seed, 32, "seed") if seed is not None else None
        full = enc_size(group, ENC_UNCOMPRESSED)
        oa, ob = np.empty(full, dtype=np.uint8), np.empty(full, dtype=np.uint8)
        fn = self.lib.p2b_g2_power_pairs if group == G2 else self.lib.p2b_g1_power_pairs
        self._check(fn(self.h, _ptr(p), n, _ptr(sc) if sc is not None else None, _ptr(sd) if sd is not None else None, scalar_bits,
                       in_enc, flags, _ptr(oa), _ptr(ob)))
        return oa.tobytes(), ob.tobytes()

    def random_scalars(self, seed, n, scalar_bits=253, first=0):
        """The coefficients the device generates for (seed, scalar_bits): n x 32 bytes big-endian."""
        out = np.empty(max(1, 32 * n), dtype=np.uint8)
        self._check(self.lib.p2b_random_scalars(self.h, _ptr(_fixed(seed, 32, "seed")), first, n, scalar_bits, _ptr(out)))
        return out[: 32 * n]

    def selftest_field(self, field, op, a, b, c=None, d=None):
        """Element-wise device field operation on raw little-endian limbs (see p2b_selftest_field); arrays of n x 32 bytes."""
        a, b = _host(a), _host(b)
        c = _host(c) if c is not None else a
        d = _host(d) if d is not None else b
        n = a.size // 32
        out = np.empty(max(1, a.size), dtype=np.uint8)
        self._check(self.lib.p2b_selftest_field(self.h, field, op, _ptr(a), _ptr(b), _ptr(c), _ptr(d), n, _ptr(out)))
        return out[: 32 * n]

    def sparse_mul(self, group, bases, row_offsets, cols, coeffs):
        """out[i] = sum_j coeffs[j] * bases[cols[j]] over CSR rows (the QAP evaluation of MPCParameters::new)."""
        b, k = _host(bases), _host(coeffs)
        ro = np.ascontiguousarray(row_offsets, dtype=np.uint64)
        cl = np.ascontiguousarray(cols, dtype=np.uint32)
        size = enc_size(group, ENC_UNCOMPRESSED)
        n_rows = ro.size - 1
        if n_rows < 0 or k.size != 32 * cl.size or (n_rows >= 0 and int(ro[-1]) != cl.size):
            raise ValueError("sparse_mul: inconsistent CSR arrays")
        out = np.empty(max(1, n_rows * size), dtype=np.uint8)
        fn = self.lib.p2b_g2_sparse_mul if group == G2 else self.lib.p2b_g1_sparse_mul
        self._check(fn(self.h, _ptr(b), b.size // size, _ptr(ro), _ptr(cl), _ptr(k), n_rows, _ptr(out)))
        return out[: n_rows * size]

    def msm_dev(self, group, d_points, d_scalars, n):
        out = np.empty(enc_size(group, ENC_UNCOMPRESSED), dtype=np.uint8)
        fn = self.lib.p2b_g2_msm_dev if group == G2 else self.lib.p2b_g1_msm_dev
        self._check(fn(self.h, _ptr(d_points), _ptr(d_scalars), n, _ptr(out)))
        return out.tobytes()

    def sum_points(self, group, points):
        """Sum of uncompressed points (combines the per-rank MSM results after the all-gather)."""
        p = _host(points)
        count = p.size // enc_size(group, ENC_UNCOMPRESSED)
        out = np.empty(enc_size(group, ENC_UNCOMPRESSED), dtype=np.uint8)
        fn = self.lib.p2b_g2_sum_points if group == G2 else self.lib.p2b_g1_sum_points
        self._check(fn(self.h, _ptr(p), count, _ptr(out)))
        return out.tobytes()

    # -- group-element FFT / prepare_phase2
    def group_fft(self, group, points, inverse=False, in_enc=ENC_UNCOMPRESSED, out_enc=ENC_UNCOMPRESSED, flags=0):
        pts = _host(points)
        d = pts.size // enc_size(group, in_enc)
        log_d = d.bit_length() - 1
        if d == 0 or (1 << log_d) != d:
            raise P2BError(EARG, "fft length must be a power of two")
        out = np.empty(d * enc_size(group, out_enc), dtype=np.uint8)
        fn = self.lib.p2b_g2_group_fft if group == G2 else self.lib.p2b_g1_group_fft
        self._check(fn(self.h, _ptr(pts), _ptr(out), log_d, int(inverse), in_enc, out_enc, flags))
        return out

    def group_fft_scaled(self, group, points, inverse, total_log_d, in_enc=ENC_UNCOMPRESSED, out_enc=ENC_UNCOMPRESSED, flags=0):
        """group_fft whose inverse scaling is 2^-total_log_d (the block-local part of a transform sharded over several GPUs)."""
        pts = _host(points)
        d = pts.size // enc_size(group, in_enc)
        log_d = d.bit_length() - 1
        if d == 0 or (1 << log_d) != d:
            raise P2BError(EARG, "fft length must be a power of two")
        out = np.empty(d * enc_size(group, out_enc), dtype=np.uint8)
        fn = self.lib.p2b_g2_group_fft_scaled if group == G2 else self.lib.p2b_g1_group_fft_scaled
        self._check(fn(self.h, _ptr(pts), _ptr(out), log_d, int(inverse), in_enc, out_enc, flags, total_log_d))
        return out

    def gfft_stage(self, group, a, b, w=None, start=0, in_enc=ENC_UNCOMPRESSED, out_enc=ENC_UNCOMPRESSED, flags=0, want_sum=True,
                   want_diff=True):
        """(a + b, [w^(start+i)] (a - b)) elementwise over two arrays of 2^k points; w=None: plain differences."""
        pa, pb = _host(a), _host(b)
        n = pa.size // enc_size(group, in_enc)
        log_n = n.bit_length() - 1
        if n == 0 or (1 << log_n) != n or pb.size != pa.size:
            raise P2BError(EARG, "gfft_stage: two arrays of the same power-of-two length")
        osz = enc_size(group, out_enc)
        s = np.empty(n * osz, dtype=np.uint8) if want_sum else None
        d = np.empty(n * osz, dtype=np.uint8) if want_diff else None
        wb = _fixed(w, 32, "w") if w is not None else None
        fn = self.lib.p2b_g2_gfft_stage if group == G2 else self.lib.p2b_g1_gfft_stage
        self._check(fn(self.h, _ptr(pa), _ptr(pb), log_n, _ptr(wb) if wb is not None else None, start, in_enc, out_enc, flags,
                       _ptr(s) if s is not None else None, _ptr(d) if d is not None else None))
        return s, d

    def pot_prepare_phase2(self, accumulator, size_log2, m, compressed_input=False, check_input=True, flags=0):
        """Image of the file phase1radix2m{m} (powersoftau/src/bin/prepare_phase2.rs:62-241)."""
        acc = _host(accumulator)
        out = np.empty(self.lib.p2b_pot_radix_file_size(m), dtype=np.uint8)
        self._check(self.lib.p2b_pot_prepare_phase2(self.h, _ptr(acc), acc.size, size_log2, int(compressed_input),
                                                    int(check_input), m, _ptr(out), out.size, flags))
        return out

    # -- FFT
    def fr_fft(self, data, inverse=False, coset=False):
        a = np.array(_host(data), dtype=np.uint8, copy=True)
        n = a.size // 32
        log_n = n.bit_length() - 1
        if n == 0 or (1 << log_n) != n:
            raise P2BError(EARG, "fft length must be a power of two")
        self._check(self.lib.p2b_fr_fft(self.h, _ptr(a), log_n, int(inverse), int(coset)))
        return a

    def fr_fft_dev(self, d_data, log_n, inverse=False, coset=False):
        self._check(self.lib.p2b_fr_fft_dev(self.h, _ptr(d_data), log_n, int(inverse), int(coset)))
