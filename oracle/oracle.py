"""ctypes binding of oracle/libp2b_oracle.so (the C restatement of the reference CPU path).

TEST INFRASTRUCTURE ONLY: imported by tests/, __graft_entry__.smoke() and bench.py's CPU-baseline
legs.  The product package never imports this module.
"""
import ctypes
import hashlib
import os
import struct
import subprocess

_HERE = os.path.dirname(os.path.abspath(__file__))
_SO = os.path.join(_HERE, "libp2b_oracle.so")

G1, G2 = 0, 1
UNCOMPRESSED, COMPRESSED, RAW_MONT_LE = 0, 1, 2
OK, EARG, EDECODE, EINFINITY_IN, EINFINITY_OUT = 0, 1, 2, 3, 4
D_NOT_ON_CURVE, D_COORD, D_UNEXPECTED_INFO, D_UNEXPECTED_COMPRESSION = 1, 2, 3, 4


def build(force=False):
    srcs = [os.path.join(_HERE, f) for f in ("p2b_oracle.c", "field.h", "ec_tmpl.h")]
    if force or not os.path.exists(_SO) or any(
            os.path.exists(s) and os.path.getmtime(s) > os.path.getmtime(_SO) for s in srcs):
        subprocess.check_call(["make", "-C", _HERE, "-B", "libp2b_oracle.so"],
                              stdout=subprocess.DEVNULL)
    return _SO


_lib = None


def lib():
    global _lib
    if _lib is None:
        if not os.path.exists(_SO):
            build()
        _lib = ctypes.CDLL(_SO)
        _lib.orc_accumulator_size.restype = ctypes.c_uint64
        _lib.orc_accumulator_size.argtypes = [ctypes.c_uint32, ctypes.c_int]
    return _lib


def point_size(group, enc):
    return (64 if group == G1 else 128) >> (1 if enc == COMPRESSED else 0)


class OracleError(Exception):
    def __init__(self, code, sub=0, index=0):
        super().__init__("oracle error code=%d sub=%d index=%d" % (code, sub, index))
        self.code, self.sub, self.index = code, sub, index


def _cbuf(b):
    return (ctypes.c_uint8 * len(b)).from_buffer_copy(b) if len(b) else (ctypes.c_uint8 * 1)()


def _cin(b):
    """Read-only input: contiguous numpy arrays are passed by pointer (no copy: the 2^26-term MSM inputs are 6.4 GB)."""
    try:
        import numpy as np
        if isinstance(b, np.ndarray) and b.flags.c_contiguous and b.dtype == np.uint8 and b.size:
            return ctypes.c_void_p(b.ctypes.data)
    except ImportError:
        pass
    return _cbuf(b)


def batch_mul(group, points, scalars, in_enc=UNCOMPRESSED, out_enc=UNCOMPRESSED, checked=False,
              reject_inf=False, threads=1):
    """scalars: bytes of n*32 (per point) or 32 (broadcast), BE canonical."""
    n = len(points) // point_size(group, in_enc)
    out = (ctypes.c_uint8 * max(1, n * point_size(group, out_enc)))()
    idx, sub = ctypes.c_uint64(0), ctypes.c_int(0)
    rc = lib().orc_batch_mul(group, _cbuf(points), out, ctypes.c_size_t(n), _cbuf(scalars),
                             ctypes.c_size_t(len(scalars) // 32), in_enc, out_enc, int(checked),
                             int(reject_inf), threads, ctypes.byref(idx), ctypes.byref(sub))
    if rc:
        raise OracleError(rc, sub.value, idx.value)
    return bytes(out)[: n * point_size(group, out_enc)]


def batch_mul_powers(group, points, tau, coeff=None, start=0, in_enc=UNCOMPRESSED,
                     out_enc=UNCOMPRESSED, checked=False, threads=1):
    n = len(points) // point_size(group, in_enc)
    out = (ctypes.c_uint8 * max(1, n * point_size(group, out_enc)))()
    idx, sub = ctypes.c_uint64(0), ctypes.c_int(0)
    rc = lib().orc_batch_mul_powers(group, _cbuf(points), out, ctypes.c_size_t(n), _cbuf(tau),
                                    _cbuf(coeff) if coeff is not None else None,
                                    ctypes.c_uint64(start), in_enc, out_enc, int(checked), threads,
                                    ctypes.byref(idx), ctypes.byref(sub))
    if rc:
        raise OracleError(rc, sub.value, idx.value)
    return bytes(out)[: n * point_size(group, out_enc)]


def accumulator_size(size_log2, compressed):
    return lib().orc_accumulator_size(size_log2, int(compressed))


def pot_generate_initial(size_log2, compressed=False):
    """Challenge file of new_constrained: blank hash + all generators (uncompressed)."""
    n = accumulator_size(size_log2, compressed)
    out = (ctypes.c_uint8 * n)()
    rc = lib().orc_pot_generate_initial(out, ctypes.c_uint64(n), size_log2, int(compressed))
    if rc:
        raise OracleError(rc)
    b = bytearray(out)
    b[0:64] = hashlib.blake2b(b"").digest()
    return bytes(b)


def pot_transform(challenge, size_log2, batch_size, tau, alpha, beta, in_compressed=False,
                  out_compressed=True, check_input=False, threads=1):
    """Returns accumulator_size(out) bytes: [0,64) = blake2b(challenge) as compute_constrained
    writes it, [64, ..) the transformed accumulator.  (The 768-byte public key is not appended.)"""
    n = accumulator_size(size_log2, out_compressed)
    out = (ctypes.c_uint8 * n)()
    idx, sub = ctypes.c_uint64(0), ctypes.c_int(0)
    rc = lib().orc_pot_transform(_cbuf(challenge), ctypes.c_uint64(len(challenge)), out,
                                 ctypes.c_uint64(n), size_log2, batch_size, int(in_compressed),
                                 int(out_compressed), int(check_input), _cbuf(tau), _cbuf(alpha),
                                 _cbuf(beta), threads, ctypes.byref(idx), ctypes.byref(sub))
    if rc:
        raise OracleError(rc, sub.value, idx.value)
    b = bytearray(out)
    b[0:64] = hashlib.blake2b(challenge).digest()
    return bytes(b)


def msm(group, points, scalars, threads=1):
    n = len(scalars) // 32
    out = (ctypes.c_uint8 * point_size(group, UNCOMPRESSED))()
    rc = lib().orc_msm(group, _cin(points), _cin(scalars), ctypes.c_size_t(n), out, threads)
    if rc:
        raise OracleError(rc)
    return bytes(out)


def msm_timed(group, points, scalars, threads=1):
    """msm() plus (wire-decode seconds, Pippenger seconds): the reference's multiexp takes decoded G1Affine / FrRepr values,
    so only the second figure is comparable with it."""
    n = len(scalars) // 32
    out = (ctypes.c_uint8 * point_size(group, UNCOMPRESSED))()
    sec = (ctypes.c_double * 2)()
    rc = lib().orc_msm_timed(group, _cin(points), _cin(scalars), ctypes.c_size_t(n), out, threads, sec)
    if rc:
        raise OracleError(rc)
    return bytes(out), sec[0], sec[1]


def sum_points(group, points):
    n = len(points) // point_size(group, UNCOMPRESSED)
    out = (ctypes.c_uint8 * point_size(group, UNCOMPRESSED))()
    rc = lib().orc_sum_points(group, _cbuf(points), ctypes.c_size_t(n), out)
    if rc:
        raise OracleError(rc)
    return bytes(out)


def sparse_mul(group, bases, row_offsets, cols, coeffs):
    """out[i] = sum_j coeffs[j] * bases[cols[j]] row by row -- the `eval` loop of MPCParameters::new
    (phase2/src/parameters.rs:244-300: `a_g1.add_assign(&coeffs_g1[lag].mul(coeff))` per (coeff, lag) entry), on the
    oracle's scalar multiplication and point sum.  Empty rows give the point at infinity."""
    size = point_size(group, UNCOMPRESSED)
    out = []
    for i in range(len(row_offsets) - 1):
        lo, hi = int(row_offsets[i]), int(row_offsets[i + 1])
        if lo == hi:
            out.append(bytes([0x40]) + bytes(size - 1))
            continue
        pts = b"".join(bases[size * int(c): size * (int(c) + 1)] for c in cols[lo:hi])
        out.append(sum_points(group, batch_mul(group, pts, bytes(coeffs[32 * lo: 32 * hi]))))
    return b"".join(out)


def fr_fft(data, inverse=False, coset=False, threads=1):
    n = len(data) // 32
    log_n = n.bit_length() - 1
    assert 1 << log_n == n
    buf = _cbuf(data)
    rc = lib().orc_fr_fft(buf, log_n, int(inverse), int(coset), threads)
    if rc:
        raise OracleError(rc)
    return bytes(buf)[: len(data)]


def point_mul(group, point, k, in_enc=UNCOMPRESSED, out_enc=UNCOMPRESSED):
    out = (ctypes.c_uint8 * point_size(group, out_enc))()
    rc = lib().orc_point_mul(group, _cbuf(point), _cbuf(k), out, in_enc, out_enc)
    if rc:
        raise OracleError(rc)
    return bytes(out)


def point_recode(group, point, in_enc, out_enc, checked=True):
    out = (ctypes.c_uint8 * point_size(group, out_enc))()
    sub = ctypes.c_int(0)
    rc = lib().orc_point_recode(group, _cbuf(point), out, in_enc, out_enc, int(checked), ctypes.byref(sub))
    if rc:
        raise OracleError(rc, sub.value)
    return bytes(out)


def field_op(field, op, a, b=bytes(32)):
    out = (ctypes.c_uint8 * 32)()
    ok = lib().orc_field_op(field, op, _cbuf(a), _cbuf(b), out)
    return bytes(out) if ok else None


def fq2_pow_raw(c0, c1, e):
    """(c0 + c1 u)^e for integers c0, c1, e: the result's raw Montgomery limbs ([c0 limbs], [c1 limbs]) as lists of 4 u64."""
    out = (ctypes.c_uint64 * 8)()
    ok = lib().orc_fq2_pow_raw(_cbuf(int(c0).to_bytes(32, "big")), _cbuf(int(c1).to_bytes(32, "big")),
                               _cbuf(int(e).to_bytes(32, "big")), out)
    assert ok
    return list(out[:4]), list(out[4:])


def fq2_mul_raw(a, b):
    """Product of two Fq2 elements given (and returned) as raw Montgomery limbs ([c0 limbs], [c1 limbs])."""
    out = (ctypes.c_uint64 * 8)()
    lib().orc_fq2_mul_raw((ctypes.c_uint64 * 8)(*(a[0] + a[1])), (ctypes.c_uint64 * 8)(*(b[0] + b[1])), out)
    return list(out[:4]), list(out[4:])


def constants():
    out = (ctypes.c_uint64 * 16)()
    lib().orc_constants(out)
    return list(out)


# ------------------------------------------------------------------ phase 2 (.params) glue
_G1_GEN = (1).to_bytes(32, "big") + (2).to_bytes(32, "big")


def _rd_vec(buf, off, size):
    (n,) = struct.unpack_from(">I", buf, off)
    off += 4
    return buf[off: off + n * size], n, off + n * size


def params_layout(buf):
    """Offsets of the sections of an MPCParameters file (bellman/src/groth16/mod.rs:252-285,
    phase2/src/parameters.rs:663-677)."""
    lay, off = {}, 0
    for name, size in (("alpha_g1", 64), ("beta_g1", 64), ("beta_g2", 128), ("gamma_g2", 128),
                       ("delta_g1", 64), ("delta_g2", 128)):
        lay[name] = (off, 1, size)
        off += size
    for name, size in (("ic", 64), ("h", 64), ("l", 64), ("a", 64), ("b_g1", 64), ("b_g2", 128)):
        (n,) = struct.unpack_from(">I", buf, off)
        lay[name] = (off + 4, n, size)
        off += 4 + n * size
    lay["cs_hash"] = (off, 1, 64)
    off += 64
    (n,) = struct.unpack_from(">I", buf, off)
    lay["contributions"] = (off + 4, n, 384)
    off += 4 + n * 384
    assert off == len(buf), (off, len(buf))
    return lay


def phase2_contribute(buf, delta, s_g1, r_g2, threads=1):
    """MPCParameters::contribute (phase2/src/parameters.rs:414-522) with the RNG-derived values
    passed explicitly: delta (32-byte BE Fr), s (64-byte G1), r (128-byte G2; hash_to_g2 of the
    transcript in the reference).  Returns (new file bytes, 64-byte contribution hash)."""
    lay = params_layout(buf)
    out = bytearray(buf)
    dint = int.from_bytes(delta, "big")
    r_mod = 21888242871839275222246405745257275088548364400416034343698204186575808495617
    dinv = pow(dint, -1, r_mod).to_bytes(32, "big")

    def sect(name):
        off, n, size = lay[name]
        return off, n * size

    for name in ("h", "l"):
        off, ln = sect(name)
        out[off: off + ln] = batch_mul(G1, bytes(buf[off: off + ln]), dinv, threads=threads)
    o1, _ = sect("delta_g1")
    o2, _ = sect("delta_g2")
    delta_after = point_mul(G1, bytes(buf[o1: o1 + 64]), delta)
    out[o1: o1 + 64] = delta_after
    out[o2: o2 + 128] = point_mul(G2, bytes(buf[o2: o2 + 128]), delta)
    s_delta = point_mul(G1, s_g1, delta)
    co, cn, _ = lay["contributions"]
    h = hashlib.blake2b()
    cs_off = lay["cs_hash"][0]
    h.update(buf[cs_off: cs_off + 64])
    h.update(buf[co: co + cn * 384])
    h.update(s_g1)
    h.update(s_delta)
    transcript = h.digest()
    r_delta = point_mul(G2, r_g2, delta)
    pubkey = delta_after + s_g1 + s_delta + r_delta + transcript
    assert len(pubkey) == 384
    out[co - 4: co] = struct.pack(">I", cn + 1)
    out += pubkey
    return bytes(out), hashlib.blake2b(pubkey).digest()
