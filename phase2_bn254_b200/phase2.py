"""Host-side mirror of the phase2 crate's MPCParameters contribution interface, backed by libp2b.so.

Mirrors  MPCParameters::{read, write, contribute}  phase2/src/parameters.rs:414-522,663-703 and the
PublicKey wire form phase2/src/keypair.rs:50-105.  Parameters are kept in their serialized form
(bellman/src/groth16/mod.rs:252-285): the GPU path consumes and produces wire bytes directly.
"""
import hashlib
import struct

import numpy as np

from . import lib as _lib


def params_layout(buf):
    """Section table of a serialized MPCParameters (bellman/src/groth16/mod.rs:252-285, phase2/src/parameters.rs:663-677):
    name -> (offset, count, element size, group).  Raises ValueError on a truncated or over-long buffer."""
    b = memoryview(buf)
    lay, off = {}, 0
    for name, size, g in (("alpha_g1", 64, 0), ("beta_g1", 64, 0), ("beta_g2", 128, 1), ("gamma_g2", 128, 1),
                          ("delta_g1", 64, 0), ("delta_g2", 128, 1)):
        lay[name] = (off, 1, size, g)
        off += size
    for name, size, g in (("ic", 64, 0), ("h", 64, 0), ("l", 64, 0), ("a", 64, 0), ("b_g1", 64, 0), ("b_g2", 128, 1)):
        if off + 4 > len(b):
            raise ValueError("params truncated (vector length)")
        (n,) = struct.unpack_from(">I", b, off)
        lay[name] = (off + 4, n, size, g)
        off += 4 + n * size
    lay["cs_hash"] = (off, 1, 64, None)
    off += 64
    if off + 4 > len(b):
        raise ValueError("params truncated (cs_hash)")
    (n,) = struct.unpack_from(">I", b, off)
    lay["contributions"] = (off + 4, n, 384, None)
    off += 4 + n * 384
    if off != len(b):
        raise ValueError("MPCParameters length mismatch: parsed %d of %d bytes" % (off, len(b)))
    return lay


class SynthesisError(Exception):
    """bellman SynthesisError (PolynomialDegreeTooLarge, UnconstrainedVariable) as raised by MPCParameters::new."""


_R = 21888242871839275222246405745257275088548364400416034343698204186575808495617


class KeypairAssembly:
    """The constraint system MPCParameters::new synthesizes the circuit into (phase2/src/keypair_assembly.rs:20-118):
    per variable, the (coefficient, constraint index) entries of the A, B and C matrices.  Variables are ("input", i) /
    ("aux", i) pairs; a linear combination is a list of (variable, coefficient) with integer coefficients mod r."""

    def __init__(self):
        self.num_inputs = self.num_aux = self.num_constraints = 0
        self.at_inputs, self.bt_inputs, self.ct_inputs = [], [], []
        self.at_aux, self.bt_aux, self.ct_aux = [], [], []

    def alloc(self):
        index = self.num_aux
        self.num_aux += 1
        self.at_aux.append([]); self.bt_aux.append([]); self.ct_aux.append([])
        return ("aux", index)

    def alloc_input(self):
        index = self.num_inputs
        self.num_inputs += 1
        self.at_inputs.append([]); self.bt_inputs.append([]); self.ct_inputs.append([])
        return ("input", index)

    def enforce(self, a, b, c):
        for lc, inputs, aux in ((a, self.at_inputs, self.at_aux), (b, self.bt_inputs, self.bt_aux),
                                (c, self.ct_inputs, self.ct_aux)):
            for (kind, index), coeff in lc:
                (inputs if kind == "input" else aux)[index].append((int(coeff) % _R, self.num_constraints))
        self.num_constraints += 1


def _csr(rows, col_shift=0):
    offs = np.zeros(len(rows) + 1, dtype=np.uint64)
    offs[1:] = np.cumsum([len(r) for r in rows], dtype=np.uint64)
    cols = np.fromiter((lag + col_shift for r in rows for _, lag in r), dtype=np.uint32, count=int(offs[-1]))
    coeffs = b"".join(int(cf).to_bytes(32, "big") for r in rows for cf, _ in r)
    return offs, cols, np.frombuffer(coeffs, dtype=np.uint8)


class MPCParameters:
    @classmethod
    def new(cls, circuit, should_filter_points_at_infinity, radix, ctx=None):
        """MPCParameters::new (phase2/src/parameters.rs:98-386).  `circuit(cs)` synthesizes into a KeypairAssembly (the
        ONE input is allocated first, the x * 0 = 0 input constraints are appended, exactly as the reference does);
        `radix(exp)` returns the bytes of phase1radix2m{exp} (prepare_phase2's output).  The QAP evaluation -- one scalar
        multiplication per matrix entry, summed per variable -- runs on the GPU (p2b_g{1,2}_sparse_mul)."""
        ctx = ctx or _lib.Context(0)
        cs = KeypairAssembly()
        cs.alloc_input()
        circuit(cs)
        for i in range(cs.num_inputs):
            cs.enforce([(("input", i), 1)], [], [])
        m, exp = 1, 0
        while m < cs.num_constraints:
            m, exp = m * 2, exp + 1
            if exp > 28:
                raise SynthesisError("PolynomialDegreeTooLarge")
        f = np.frombuffer(bytes(radix(exp)), dtype=np.uint8)
        if f.size < 192 + 384 * m:
            raise IOError("UnexpectedEof: phase1radix2m%d is too short" % exp)
        sizes = (("alpha", 64, 1), ("beta_g1", 64, 1), ("beta_g2", 128, 1), ("coeffs_g1", 64, m), ("coeffs_g2", 128, m),
                 ("alpha_coeffs_g1", 64, m), ("beta_coeffs_g1", 64, m), ("h", 64, m - 1))
        sec, off = {}, 0
        for name, size, cnt in sizes:
            sec[name] = f[off: off + size * cnt]
            if cnt and (sec[name].reshape(cnt, size)[:, 0] & 0x40).any():
                raise IOError("InvalidData: point at infinity")                      # read_g1 / read_g2 (:156-180)
            off += size * cnt
        at, bt, ct = cs.at_inputs + cs.at_aux, cs.bt_inputs + cs.bt_aux, cs.ct_inputs + cs.ct_aux
        a_off, a_cols, a_k = _csr(at)
        b_off, b_cols, b_k = _csr(bt)
        a_g1 = ctx.sparse_mul(0, sec["coeffs_g1"], a_off, a_cols, a_k)
        b_g1 = ctx.sparse_mul(0, sec["coeffs_g1"], b_off, b_cols, b_k)
        b_g2 = ctx.sparse_mul(1, sec["coeffs_g2"], b_off, b_cols, b_k)
        # ext = A . beta_coeffs + B . alpha_coeffs + C . coeffs  (:285-297): one product over the three bases side by side
        ext_rows = [[(cf, lag) for cf, lag in ra] + [(cf, lag + m) for cf, lag in rb] + [(cf, lag + 2 * m) for cf, lag in rc]
                    for ra, rb, rc in zip(at, bt, ct)]
        e_off, e_cols, e_k = _csr(ext_rows)
        ext = ctx.sparse_mul(0, np.concatenate([sec["beta_coeffs_g1"], sec["alpha_coeffs_g1"], sec["coeffs_g1"]]),
                             e_off, e_cols, e_k)
        ic, l = ext[: 64 * cs.num_inputs], ext[64 * cs.num_inputs:]
        if l.size and (l.reshape(-1, 64)[:, 0] & 0x40).any():
            raise SynthesisError("UnconstrainedVariable")

        def vec(points, size, filt):
            p = points.reshape(-1, size)
            if filt:
                p = p[(p[:, 0] & 0x40) == 0]
            return struct.pack(">I", p.shape[0]) + p.tobytes()

        from .powersoftau import G1_ONE, G2_ONE
        flt = bool(should_filter_points_at_infinity)
        body = (sec["alpha"].tobytes() + sec["beta_g1"].tobytes() + sec["beta_g2"].tobytes() + G2_ONE + G1_ONE + G2_ONE +
                vec(ic, 64, False) + vec(sec["h"], 64, False) + vec(l, 64, False) + vec(a_g1, 64, flt) +
                vec(b_g1, 64, flt) + vec(b_g2, 128, flt))
        cs_hash = hashlib.blake2b(body).digest()                                     # HashWriter over params.write (:365-375)
        return cls(body + cs_hash + struct.pack(">I", 0))

    def __init__(self, data, validated=False):
        self.data = np.array(np.frombuffer(bytes(data), dtype=np.uint8)) if not isinstance(data, np.ndarray) else data
        # True once every point went through the checked codec (read(checked=True)): the MSM kernels decode unchecked, so
        # verify_contribution re-validates H / L of an object that was constructed from raw bytes
        self.validated = bool(validated)

    @classmethod
    def read(cls, reader, disallow_points_at_infinity=False, checked=True, ctx=None):
        """MPCParameters::read (phase2/src/parameters.rs:682-703 -> Parameters::read, groth16/mod.rs:287-383).
        As in the reference the caller's flags only govern h, l, a, b_g1, b_g2 (`read_g1` / `read_g2`, mod.rs:295-327).
        The verifying key is always curve-checked (`into_affine()`, mod.rs:160-185) and its `ic` always rejects infinity
        (mod.rs:188-199); every stored contribution's delta_after, s, s_delta (G1) and r_delta (G2) is always curve-checked
        and must not be infinity (PublicKey::read, phase2/src/keypair.rs:64-100).  All of it runs section by section on
        the GPU bulk codec.  Raises IOError("InvalidData ...") like the reference's io::Error."""
        buf = reader if isinstance(reader, (bytes, bytearray, memoryview, np.ndarray)) else reader.read()
        self = cls(buf)
        ctx = ctx or _lib.Context(0)
        caller = (_lib.CHECK_INPUT if checked else 0) | (_lib.REJECT_INFINITY if disallow_points_at_infinity else 0)
        lay = params_layout(self.data)

        def validate(name, group, points, flags, index_of=lambda i: i):
            if points.size == 0:
                return
            try:
                ctx.recode(group, points, _lib.ENC_UNCOMPRESSED, _lib.ENC_UNCOMPRESSED, flags)
            except _lib.P2BError as e:
                if e.code in (_lib.EDECODE, _lib.EINFINITY_IN):
                    what = "point at infinity" if e.code == _lib.EINFINITY_IN else \
                        {1: "NotOnCurve", 2: "CoordinateDecodingError", 3: "UnexpectedInformation",
                         4: "UnexpectedCompressionMode"}.get(e.sub, "decoding error")
                    raise IOError("InvalidData: %s in %s[%d]" % (what, name, index_of(e.index)))
                raise

        for name, (off, n, size, group) in lay.items():
            if group is None or n == 0:
                continue
            if name in ("alpha_g1", "beta_g1", "beta_g2", "gamma_g2", "delta_g1", "delta_g2"):
                flags = _lib.CHECK_INPUT
            elif name == "ic":
                flags = _lib.CHECK_INPUT | _lib.REJECT_INFINITY
            else:
                flags = caller
            validate(name, group, self.data[off: off + n * size], flags)
        off, n, _, _ = lay["contributions"]
        if n:
            pk = self.data[off: off + n * 384].reshape(n, 384)
            strict = _lib.CHECK_INPUT | _lib.REJECT_INFINITY
            validate("contributions(delta_after, s, s_delta)", 0, np.ascontiguousarray(pk[:, :192]).reshape(-1), strict,
                     lambda i: i // 3)
            validate("contributions(r_delta)", 1, np.ascontiguousarray(pk[:, 192:320]).reshape(-1), strict)
        self.validated = bool(checked)
        return self

    def write(self, writer):
        writer.write(self.data.tobytes())

    def section(self, name):
        """The wire bytes of one section (numpy view)."""
        off, n, size, _ = params_layout(self.data)[name]
        return self.data[off: off + n * size]

    def contribute(self, delta=None, s_g1=None, r_g2=None, ctx=None, hash_to_g2=None, rng=None):
        """Contributes `delta` (int in [1, r)).  The reference draws delta, s = G1::rand and r = hash_to_g2(transcript)
        from its ChaCha RNG (parameters.rs:860-908): pass `rng` (a lib.ChaChaRng) to do the same, or make them explicit:
        delta and s_g1 plus r_g2, or a `hash_to_g2` callable mapping the 64-byte transcript to a 128-byte uncompressed
        G2 point.  Returns the 64-byte contribution hash."""
        ctx = ctx or _lib.Context(0)
        if rng is not None:
            delta, s_g1 = rng.gen_fr(), np.frombuffer(rng.gen_g1(), dtype=np.uint8)      # keypair(): delta, then s
            hash_to_g2 = hash_to_g2 or _lib.hash_to_g2
        d = np.frombuffer(int(delta).to_bytes(32, "big"), dtype=np.uint8)
        if r_g2 is None:
            if hash_to_g2 is None:
                raise ValueError("need r_g2 or hash_to_g2")
            r_g2 = hash_to_g2(ctx.phase2_transcript(self.data, d, s_g1))
        out, h = ctx.phase2_contribute(self.data, d, s_g1, r_g2)
        self.data = out
        return h


class VerificationError(Exception):
    """The `Err(())` of verify_contribution / MPCParameters::verify, with the failed check named."""


def merge_pairs(ctx, v1, v2, rng=None, scalar_bits=253, flags=0):
    """Random linear combination (sum rho_i v1_i, sum rho_i v2_i) over G1 vectors (phase2/src/utils.rs:59-105) in one pass on
    the GPU; the coefficients are generated on the device from 32 bytes of `rng` (the OS CSPRNG by default).
    `scalar_bits`: see powersoftau._random_scalars (253 = full-size scalars like the reference's Fr::rand)."""
    from .powersoftau import merge_pairs as _merge
    n = v1.size // 64
    if n != v2.size // 64:
        raise ValueError("merge_pairs: length mismatch")
    if n == 0:
        zero = bytes([0x40]) + bytes(63)
        return zero, zero
    return _merge(ctx, 0, v1, v2, None, rng, scalar_bits, flags=flags)


def verify_contribution(before, after, ctx=None, rng=None, scalar_bits=253, merge=None):
    """verify_contribution(before, after) -> the 64-byte hash of the new contribution (parameters.rs:722-855).
    Raises VerificationError where the reference returns Err(()).  `merge(v1, v2) -> (s, sx)` replaces the local merge_pairs
    (dist.sharded_verify_contribution passes the multi-GPU one)."""
    lb, la = params_layout(before.data), params_layout(after.data)
    raw = lambda m, lay, name: m.data[lay[name][0]: lay[name][0] + lay[name][1] * lay[name][2]]

    def fail(what):
        raise VerificationError(what)

    nb, na = lb["contributions"][1], la["contributions"][1]
    if na != nb + 1:
        fail("transformation must add exactly one contribution")
    if raw(before, lb, "contributions").tobytes() != raw(after, la, "contributions")[: nb * 384].tobytes():
        fail("previous contributions changed")
    for name in ("h", "l"):
        if lb[name][1] != la[name][1]:
            fail("%s changed length" % name)
    for name in ("a", "b_g1", "b_g2", "alpha_g1", "beta_g1", "beta_g2", "gamma_g2", "ic", "cs_hash"):
        if lb[name][1] != la[name][1] or raw(before, lb, name).tobytes() != raw(after, la, name).tobytes():
            fail("%s changed" % name)
    pk = raw(after, la, "contributions")[nb * 384:].tobytes()
    delta_after, s, s_delta, r_delta, transcript = pk[:64], pk[64:128], pk[128:192], pk[192:320], pk[320:384]
    h = hashlib.blake2b()
    h.update(raw(before, lb, "cs_hash").tobytes())
    h.update(raw(before, lb, "contributions").tobytes())
    h.update(s)
    h.update(s_delta)
    if transcript != h.digest():
        fail("transcript is inconsistent")
    r = _lib.hash_to_g2(h.digest())
    try:
        if not _lib.same_ratio((s, s_delta), (r, r_delta)):                        # same_ratio((r, r_delta), (s, s_delta))
            fail("signature of knowledge")
        d1b, d1a = raw(before, lb, "delta_g1").tobytes(), raw(after, la, "delta_g1").tobytes()
        d2b, d2a = raw(before, lb, "delta_g2").tobytes(), raw(after, la, "delta_g2").tobytes()
        if not _lib.same_ratio((d1b, delta_after), (r, r_delta)):
            fail("delta_g1 is not the old delta times the new one")
        if delta_after != d1a:
            fail("delta_after != delta_g1")
        g1_one = (1).to_bytes(32, "big") + (2).to_bytes(32, "big")
        from .powersoftau import G2_ONE
        if not _lib.same_ratio((g1_one, delta_after), (G2_ONE, d2a)):
            fail("delta_g2 is inconsistent with delta_g1")
        # The MSM decodes its points unchecked unless asked.  The reference can only get here through
        # MPCParameters::read(checked = true) (verify.rs); for objects built from raw bytes the curve check of H / L rides along
        # in the same pass (P2B_CHECK_INPUT in the MSM's decode kernel).
        unvalidated = not (getattr(before, "validated", False) and getattr(after, "validated", False))
        if merge is None:
            ctx = ctx or _lib.Context(0)                                            # only the H / L checks need the GPU
            merge = lambda a, b: merge_pairs(ctx, a, b, rng, scalar_bits, _lib.CHECK_INPUT if unvalidated else 0)
        elif unvalidated:
            ctx = ctx or _lib.Context(0)
            for m, lay in ((before, lb), (after, la)):
                if not getattr(m, "validated", False):
                    for name in ("h", "l"):
                        if lay[name][1]:
                            ctx.validate(0, raw(m, lay, name), _lib.ENC_UNCOMPRESSED, _lib.CHECK_INPUT)
        for name in ("h", "l"):                                                     # updated with delta^-1: ratios reversed
            if not _lib.same_ratio(merge(raw(before, lb, name), raw(after, la, name)), (d2a, d2b)):
                fail("%s query was not multiplied by delta^-1" % name)
    except _lib.P2BError as e:
        fail("a point does not decode: %s" % e)
    return hashlib.blake2b(pk).digest()


def keypair(rng, current):
    """keypair(rng, current) -> (public key bytes [384], delta) (parameters.rs:860-908)."""
    lay = params_layout(current.data)
    raw = lambda name: current.data[lay[name][0]: lay[name][0] + lay[name][1] * lay[name][2]].tobytes()
    delta = rng.gen_fr()
    db = int(delta).to_bytes(32, "big")
    s = rng.gen_g1()
    s_delta = _lib.host_mul(0, s, db)
    h = hashlib.blake2b(raw("cs_hash") + raw("contributions") + s + s_delta).digest()
    r = _lib.hash_to_g2(h)
    r_delta = _lib.host_mul(1, r, db)
    return _lib.host_mul(0, raw("delta_g1"), db) + s + s_delta + r_delta + h, delta
