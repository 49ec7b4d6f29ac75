"""Host-side mirror of the powersoftau crate's contribution interface, backed by libp2b.so.

Mirrors (names, argument meaning, error behaviour):
  CeremonyParams         powersoftau/src/parameters.rs:38-120
  UseCompression, CheckForCorrectness   parameters.rs:127-140
  DeserializationError   parameters.rs:143-170
  PrivateKey             powersoftau/src/keypair.rs:47-51
  BatchedAccumulator.transform   powersoftau/src/batched_accumulator.rs:1119-1292
  BatchedAccumulator.verify_transformation   batched_accumulator.rs:279-540 (MSMs on the GPU, pairings on the host)
  PublicKey, keypair, compute_g2_s, hash_to_g2, same_ratio   keypair.rs:29-213, utils.rs:31-45,151-185
The maps are numpy uint8 arrays (np.memmap works), exactly the byte layout of the challenge / response files.
"""
import hashlib
import os
from concurrent.futures import ThreadPoolExecutor
from dataclasses import dataclass

import numpy as np

from . import lib as _lib


_FQ = 21888242871839275222246405745257275088696311157297823662689037894645226208583


class UseCompression:
    Yes, No = True, False


class CheckForCorrectness:
    Yes, No = True, False


class DeserializationError(Exception):
    """IoError / DecodingError(GroupDecodingError) / PointAtInfinity."""

    def __init__(self, kind, detail=""):
        super().__init__("%s %s" % (kind, detail))
        self.kind = kind


class CeremonyParams:
    """Bn256 geometry: g1 64/32 bytes, g2 128/64 bytes (parameters.rs:21-34,72-120)."""

    def __init__(self, size, batch_size):
        self.size, self.batch_size = size, batch_size
        self.g1, self.g2, self.g1_compressed, self.g2_compressed = 64, 128, 32, 64
        self.powers_length = 1 << size
        self.powers_g1_length = (self.powers_length << 1) - 1
        self.hash_size = 64
        self.accumulator_size = (self.powers_g1_length * self.g1 + self.powers_length * self.g2 +
                                 self.powers_length * self.g1 + self.powers_length * self.g1 + self.g2 +
                                 self.hash_size)
        self.public_key_size = 3 * self.g2 + 6 * self.g1
        self.contribution_size = (self.powers_g1_length * self.g1_compressed +
                                  self.powers_length * self.g2_compressed +
                                  self.powers_length * self.g1_compressed +
                                  self.powers_length * self.g1_compressed + self.g2_compressed + self.hash_size +
                                  self.public_key_size)


@dataclass
class PrivateKey:
    """tau, alpha, beta as integers in [0, r) (keypair.rs:47-51)."""
    tau: int
    alpha: int
    beta: int


G1_ONE = (1).to_bytes(32, "big") + (2).to_bytes(32, "big")                          # ec.rs:1013-1051
G2_ONE = b"".join(v.to_bytes(32, "big") for v in (                                  # fq.rs:54-83 (c1 first)
    11559732032986387107991004021392285783925812861821192530917403151452391805634,
    10857046999023057135944570762232829481370756359578518086990519993285655852781,
    4082367875863433681332203403145435568316851327593401208105741076214120093531,
    8495653923123431417604973247489272438418190587263600148770280649306958101930))

same_ratio = _lib.same_ratio          # utils.rs:151-159
hash_to_g2 = _lib.hash_to_g2          # utils.rs:31-45


def compute_g2_s(digest, g1_s, g1_s_x, personalization):
    """Blake2b(personalization | digest | g1_s | g1_s_x) hashed into G2 (utils.rs:172-185)."""
    h = hashlib.blake2b()
    h.update(bytes([personalization]))
    h.update(bytes(digest))
    h.update(bytes(g1_s))
    h.update(bytes(g1_s_x))
    return hash_to_g2(h.digest())


@dataclass
class PublicKey:
    """keypair.rs:29-45: (g1_s, g1_s_x) pairs and g2_s_x for tau, alpha, beta; all uncompressed wire points."""
    tau_g1: tuple
    alpha_g1: tuple
    beta_g1: tuple
    tau_g2: bytes
    alpha_g2: bytes
    beta_g2: bytes

    def serialize(self):
        """768 bytes, always uncompressed (keypair.rs:105-122)."""
        return b"".join(bytes(x) for x in (self.tau_g1[0], self.tau_g1[1], self.alpha_g1[0], self.alpha_g1[1],
                                           self.beta_g1[0], self.beta_g1[1], self.tau_g2, self.alpha_g2, self.beta_g2))

    @classmethod
    def deserialize(cls, data):
        """Always checked, no points at infinity (keypair.rs:127-167)."""
        b = bytes(data)
        if len(b) < 768:
            raise DeserializationError("IoError", "public key is 768 bytes")
        one = (1).to_bytes(32, "big")
        pts = [b[64 * i: 64 * i + 64] for i in range(6)] + [b[384 + 128 * i: 512 + 128 * i] for i in range(3)]
        for i, pt in enumerate(pts):
            try:
                if _lib.host_mul(int(i >= 6), pt, one)[0] & 0x40:
                    raise DeserializationError("PointAtInfinity", "in the public key")
            except _lib.P2BError:
                raise DeserializationError("DecodingError", "public key element %d" % i)
        return cls((pts[0], pts[1]), (pts[2], pts[3]), (pts[4], pts[5]), pts[6], pts[7], pts[8])

    def write(self, output_map, accumulator_was_compressed, parameters):
        """At the end of the response map (keypair.rs:170-191)."""
        size = parameters.contribution_size if accumulator_was_compressed else parameters.accumulator_size + 768
        pos = size - parameters.public_key_size if accumulator_was_compressed else parameters.accumulator_size
        output_map[pos: pos + 768] = np.frombuffer(self.serialize(), dtype=np.uint8)

    @classmethod
    def read(cls, input_map, accumulator_was_compressed, parameters):
        """keypair.rs:193-213."""
        pos = (parameters.contribution_size - parameters.public_key_size if accumulator_was_compressed
               else parameters.accumulator_size)
        return cls.deserialize(np.asarray(input_map[pos: pos + 768]).tobytes())


def keypair(rng, digest):
    """keypair(rng, digest) -> (PublicKey, PrivateKey) (keypair.rs:54-103).  `rng` is a lib.ChaChaRng (the reference
    seeds one from OS entropy + user input, compute_constrained.rs:103-141, or from the beacon hash)."""
    assert len(digest) == 64
    tau, alpha, beta = rng.gen_fr(), rng.gen_fr(), rng.gen_fr()
    return public_key_for(PrivateKey(tau, alpha, beta), rng, digest), PrivateKey(tau, alpha, beta)


def public_key_for(key, rng, digest):
    """The PublicKey half of keypair() for given secrets: three proofs of knowledge (g1_s, g1_s^x, g2_s^x) with g1_s drawn
    from `rng` and g2_s = hash of (personalization, digest, g1_s, g1_s^x) into G2 (keypair.rs:64-90)."""
    tau, alpha, beta = key.tau, key.alpha, key.beta

    def op(x, personalization):
        g1_s = rng.gen_g1()
        xb = int(x).to_bytes(32, "big")
        g1_s_x = _lib.host_mul(0, g1_s, xb)
        g2_s = compute_g2_s(digest, g1_s, g1_s_x, personalization)
        return (g1_s, g1_s_x), _lib.host_mul(1, g2_s, xb)

    pk_tau, pk_alpha, pk_beta = op(tau, 0), op(alpha, 1), op(beta, 2)
    return PublicKey(pk_tau[0], pk_alpha[0], pk_beta[0], pk_tau[1], pk_alpha[1], pk_beta[1])


def calculate_hash(input_map):
    """Blake2b-512 over the whole map (powersoftau/src/utils.rs:20-27)."""
    h = hashlib.blake2b()
    a = np.asarray(input_map, dtype=np.uint8).reshape(-1)
    step = 1 << 26
    for off in range(0, a.size, step):
        h.update(a[off: off + step].data)
    return h.digest()


class BatchedAccumulator:
    """Only `transform` lives on the GPU path; the class is a namespace like the reference's impl block."""

    _ctx = None

    @classmethod
    def context(cls, device=0):
        if cls._ctx is None or cls._ctx.device != device:
            cls._ctx = _lib.Context(device)
        return cls._ctx

    @staticmethod
    def generate_initial(output_map, compress_the_output, parameters):
        """The initial accumulator of new_constrained: every element is the group generator
        (batched_accumulator.rs:1295-1347).  Writes output_map[64:]; the blank-hash prefix is the caller's
        (new_constrained.rs:56-63).  Pure byte replication -- no arithmetic, runs on the host."""
        g1 = (1).to_bytes(32, "big") + (2).to_bytes(32, "big")                      # G1 generator (1, 2), ec.rs:1013-1051
        g2 = b"".join(v.to_bytes(32, "big") for v in (                              # G2 generator, fq.rs:54-83 (c1 first)
            11559732032986387107991004021392285783925812861821192530917403151452391805634,
            10857046999023057135944570762232829481370756359578518086990519993285655852781,
            4082367875863433681332203403145435568316851327593401208105741076214120093531,
            8495653923123431417604973247489272438418190587263600148770280649306958101930))
        if compress_the_output:
            g1c = bytearray(g1[:32])
            if 2 > _FQ - 2:
                g1c[0] |= 0x80
            y1, y0 = int.from_bytes(g2[64:96], "big"), int.from_bytes(g2[96:128], "big")
            larger = (y1 > _FQ - y1) if y1 else (y0 > _FQ - y0)                      # Fq2 order: c1 then c0 (fq2.rs:21-30)
            g2c = bytearray(g2[:64])
            if larger:
                g2c[0] |= 0x80
            g1, g2 = bytes(g1c), bytes(g2c)
        a1, a2 = np.frombuffer(g1, dtype=np.uint8), np.frombuffer(g2, dtype=np.uint8)
        p = parameters
        o = 64
        for cnt, g in ((p.powers_g1_length, a1), (p.powers_length, a2), (p.powers_length, a1), (p.powers_length, a1), (1, a2)):
            output_map[o:o + cnt * g.size].reshape(cnt, g.size)[:] = g
            o += cnt * g.size

    @classmethod
    def decompress(cls, input_map, output_map, check_input_for_correctness, parameters, ctx=None, shard_index=0,
                   shard_count=1):
        """Compressed response -> uncompressed accumulator (batched_accumulator.rs:543-618).  Writes
        output_map[64 : accumulator_size]; the hash prefix is the caller's (verify_transform_constrained.rs:207-229)."""
        ctx = ctx or cls.context()
        try:
            ctx.pot_decompress(input_map, output_map, parameters.size, bool(check_input_for_correctness), shard_index,
                               shard_count)
        except _lib.P2BError as e:
            cls._raise(e)

    @staticmethod
    def _raise(e):
        if e.code == _lib.EDECODE:
            names = {1: "NotOnCurve", 2: "CoordinateDecodingError", 3: "UnexpectedInformation",
                     4: "UnexpectedCompressionMode"}
            raise DeserializationError("DecodingError", "%s at element %d" % (names.get(e.sub, "?"), e.index))
        if e.code == _lib.EINFINITY_IN:
            raise DeserializationError("PointAtInfinity", "at element %d" % e.index)
        if e.code == _lib.EINFINITY_OUT:
            raise AssertionError("your contribution happened to produce a point at infinity, please re-run")
        raise e

    @classmethod
    def transform(cls, input_map, output_map, input_is_compressed, compress_the_output,
                  check_input_for_correctness, key, parameters, ctx=None, shard_index=0, shard_count=1,
                  g2_in_subgroup=False):
        """Transforms the accumulator with a private key (batched_accumulator.rs:1119-1292).

        Writes output_map[64 : accumulator end]; bytes [0, 64) and the public key tail are the caller's, as in
        compute_constrained.rs:155-161,207-209.  Raises DeserializationError where the reference's
        read_chunk(...).expect() panics and AssertionError where it asserts on a produced point at infinity."""
        ctx = ctx or cls.context()
        need_in = ctx.pot_accumulator_size(parameters.size, input_is_compressed)
        if len(input_map) < need_in:
            raise ValueError("The size of challenge file should be %d, but it's %d" % (need_in, len(input_map)))
        be = lambda v: np.frombuffer(int(v).to_bytes(32, "big"), dtype=np.uint8)
        try:
            ctx.pot_transform(input_map, output_map, parameters.size, parameters.batch_size, be(key.tau),
                              be(key.alpha), be(key.beta), bool(input_is_compressed), bool(compress_the_output),
                              int(bool(check_input_for_correctness)) | (_lib.G2_SUBGROUP if g2_in_subgroup else 0),
                              shard_index, shard_count)
        except _lib.P2BError as e:
            if e.code == _lib.EDECODE:
                names = {1: "NotOnCurve", 2: "CoordinateDecodingError", 3: "UnexpectedInformation",
                         4: "UnexpectedCompressionMode"}
                raise DeserializationError("DecodingError", "%s at element %d" % (names.get(e.sub, "?"), e.index))
            if e.code == _lib.EINFINITY_IN:
                raise DeserializationError("PointAtInfinity", "at element %d" % e.index)
            if e.code == _lib.EINFINITY_OUT:
                raise AssertionError("your contribution happened to produce a point at infinity, please re-run")
            raise


def contribute_challenge(input_map, output_map, rng, parameters, input_is_compressed=False, compress_the_output=True,
                         check_input_for_correctness=False, ctx=None, overlap=True, g2_in_subgroup=False):
    """One participant's step, the body of the `compute_constrained` / `beacon_constrained` binaries
    (powersoftau/src/bin/compute_constrained.rs:140-230): hash the challenge, write that hash to the head of the response, draw
    the key pair, transform, append the public key, hash the response.  Returns (challenge_hash, response_hash, PublicKey).

    The two BLAKE2b-512 file hashes are sequential chains at about 1 GB/s on one host core (utils.rs:20-27) -- at 2^20 powers the
    402 MB challenge alone takes longer to hash than the GPU takes to transform it.  Only the PUBLIC key depends on the challenge
    hash (keypair.rs:64-90), the secrets come first out of the generator (keypair.rs:58-62), so with `overlap` the challenge is
    hashed on a host thread while the GPU transforms (hashlib and the C ABI call both release the interpreter lock); the
    generator is consumed in the reference's order either way, and the bytes written are the same (tests/test_verify_host.py).
    The response hash needs the finished file, head included, and stays last."""
    import threading
    tau, alpha, beta = rng.gen_fr(), rng.gen_fr(), rng.gen_fr()
    key = PrivateKey(tau, alpha, beta)
    box = {}
    hasher = threading.Thread(target=lambda: box.setdefault("h", calculate_hash(input_map)))
    hasher.start()
    if not overlap:
        hasher.join()
    try:
        BatchedAccumulator.transform(input_map, output_map, input_is_compressed, compress_the_output, check_input_for_correctness,
                                     key, parameters, ctx=ctx, g2_in_subgroup=g2_in_subgroup)
    finally:
        hasher.join()
    digest = box["h"]
    output_map[:64] = np.frombuffer(digest, dtype=np.uint8)
    pubkey = public_key_for(key, rng, digest)
    pubkey.write(output_map, compress_the_output, parameters)
    return digest, calculate_hash(output_map), pubkey


def _sections(parameters, compressed):
    """Byte offset and element size of the five accumulator sections (batched_accumulator.rs:96-178)."""
    p = parameters
    g1, g2 = (p.g1_compressed, p.g2_compressed) if compressed else (p.g1, p.g2)
    out, off = {}, p.hash_size
    for name, cnt, size, group in (("tau_g1", p.powers_g1_length, g1, 0), ("tau_g2", p.powers_length, g2, 1),
                                   ("alpha_g1", p.powers_length, g1, 0), ("beta_g1", p.powers_length, g1, 0),
                                   ("beta_g2", 1, g2, 1)):
        out[name] = (off, cnt, size, group)
        off += cnt * size
    return out


def _read_points(ctx, amap, parameters, compressed, checked, name, start, count):
    """read_chunk for one element type (batched_accumulator.rs:889-1001): `count` points from index `start` as
    uncompressed wire bytes; decompression and the optional curve check run on the GPU codec; infinity is rejected."""
    off, total, size, group = _sections(parameters, compressed)[name]
    count = max(0, min(count, total - start))
    raw = np.asarray(amap[off + start * size: off + (start + count) * size])
    try:
        return ctx.recode(group, raw, _lib.ENC_COMPRESSED if compressed else _lib.ENC_UNCOMPRESSED, _lib.ENC_UNCOMPRESSED,
                          (_lib.CHECK_INPUT if checked else 0) | _lib.REJECT_INFINITY)
    except _lib.P2BError as e:
        BatchedAccumulator._raise(e)


class system_rng:
    """Coefficients for the verifier's random linear combinations from the operating system's CSPRNG (`os.urandom`) -- the
    counterpart of the reference's `thread_rng()` (powersoftau/src/utils.rs:118-124, phase2/src/utils.rs:76-78).  An
    adversarial contributor must not be able to predict them, so a seedable statistical generator is not a valid default;
    tests and benches may still pass any object with a `bytes(n)` method (e.g. numpy's Generator) for reproducibility."""

    @staticmethod
    def bytes(n):
        return os.urandom(n)


MIN_SCALAR_BITS = 128


def _random_scalars(rng, n, bits=253):
    """n random scalars below 2^bits (32 bytes big-endian each).  The reference draws full-size Fr::rand from thread_rng
    (utils.rs:118-124); 253 bits (< r) is the faithful default.  bits=128 is an opt-in for big verifications: the soundness
    error of the random linear combination is 2^-128 instead of 2^-253, the zero top digits never reach a bucket (the MSM
    does about half the work) and half the random bytes are drawn.  Fewer than 128 bits are refused."""
    if not MIN_SCALAR_BITS <= bits <= 253 or bits % 8 not in (0, 5):
        raise ValueError("scalar_bits must be a multiple of 8 in [%d, 248], or 253" % MIN_SCALAR_BITS)
    nbytes = (bits + 7) // 8
    a = np.zeros((n, 32), dtype=np.uint8)
    a[:, 32 - nbytes:] = np.frombuffer(rng.bytes(nbytes * n), dtype=np.uint8).reshape(n, nbytes)
    if bits % 8:
        a[:, 32 - nbytes] &= (1 << (bits % 8)) - 1
    return a.reshape(-1)


def verify_transformation(input_map, output_map, key, digest, input_is_compressed, output_is_compressed,
                          check_input_for_correctness, check_output_for_correctness, parameters, ctx=None, rng=None,
                          scalar_bits=253):
    """BatchedAccumulator::verify_transformation (batched_accumulator.rs:279-540): the proofs of knowledge, the ratio
    checks on the first elements, then chunk by chunk `same_ratio(power_pairs(after.*), (tau_g2[0], tau_g2[1]))`.
    power_pairs = two Pippenger MSMs per element type and chunk on the GPU; same_ratio = two pairings on the host.
    Returns True / False like the reference; a chunk that does not deserialize raises (the reference panics).
    `scalar_bits`: size of the random coefficients (see _random_scalars; 253 = the reference's full-size scalars)."""
    ctx = ctx or BatchedAccumulator.context()
    _random_scalars(system_rng(), 0, scalar_bits)       # refuse a too-small scalar_bits before any work
    assert len(digest) == 64
    p = parameters
    # The per-chunk same_ratio checks (two pairings each, ~35 ms on one core) run on host threads while the GPU works on
    # the next chunk's MSMs (the library call releases the GIL); the verdict is the conjunction, as in the reference.
    pool = ThreadPoolExecutor(max_workers=max(1, min(32, os.cpu_count() or 1)))
    try:
        return _verify_transformation(pool, ctx, rng, input_map, output_map, key, digest, input_is_compressed,
                                      output_is_compressed, check_input_for_correctness, check_output_for_correctness, p,
                                      scalar_bits)
    finally:
        pool.shutdown(wait=False, cancel_futures=True)


def _verify_transformation(pool, ctx, rng, input_map, output_map, key, digest, input_is_compressed, output_is_compressed,
                           check_input_for_correctness, check_output_for_correctness, p, scalar_bits):
    tau_g2_s = compute_g2_s(digest, key.tau_g1[0], key.tau_g1[1], 0)
    alpha_g2_s = compute_g2_s(digest, key.alpha_g1[0], key.alpha_g1[1], 1)
    beta_g2_s = compute_g2_s(digest, key.beta_g1[0], key.beta_g1[1], 2)
    # proofs of knowledge: g1^s / g1^(s*x) = g2^s / g2^(s*x)
    if not same_ratio(key.tau_g1, (tau_g2_s, key.tau_g2)):
        return False
    if not same_ratio(key.alpha_g1, (alpha_g2_s, key.alpha_g2)):
        return False
    if not same_ratio(key.beta_g1, (beta_g2_s, key.beta_g2)):
        return False

    def rd(amap, compressed, checked, name, start, count):
        return _read_points(ctx, amap, p, compressed, checked, name, start, count)

    before = lambda name, start, count: rd(input_map, input_is_compressed, check_input_for_correctness, name, start, count)
    after = lambda name, start, count: rd(output_map, output_is_compressed, check_output_for_correctness, name, start, count)
    pt = lambda a, i, size: a[i * size: (i + 1) * size].tobytes()

    b_tau, a_tau = before("tau_g1", 0, 2), after("tau_g1", 0, 2)
    a_tau2 = after("tau_g2", 0, 2)
    b_alpha, a_alpha = before("alpha_g1", 0, 2), after("alpha_g1", 0, 2)
    b_beta, a_beta = before("beta_g1", 0, 2), after("beta_g1", 0, 2)
    b_beta2, a_beta2 = before("beta_g2", 0, 1), after("beta_g2", 0, 1)
    if pt(a_tau, 0, 64) != G1_ONE or pt(a_tau2, 0, 128) != G2_ONE:
        return False
    if not same_ratio((pt(b_tau, 1, 64), pt(a_tau, 1, 64)), (tau_g2_s, key.tau_g2)):
        return False
    if not same_ratio((pt(b_alpha, 0, 64), pt(a_alpha, 0, 64)), (alpha_g2_s, key.alpha_g2)):
        return False
    if not same_ratio((pt(b_beta, 0, 64), pt(a_beta, 0, 64)), (beta_g2_s, key.beta_g2)):
        return False
    if not same_ratio((pt(b_beta, 0, 64), pt(a_beta, 0, 64)), (pt(b_beta2, 0, 128), pt(a_beta2, 0, 128))):
        return False
    g2_pair = (pt(a_tau2, 0, 128), pt(a_tau2, 1, 128))
    g1_pair = (pt(a_tau, 0, 64), pt(a_tau, 1, 64))

    pending = []
    sec_in, sec_out = _sections(p, input_is_compressed), _sections(p, output_is_compressed)
    enc_in = _lib.ENC_COMPRESSED if input_is_compressed else _lib.ENC_UNCOMPRESSED
    enc_out = _lib.ENC_COMPRESSED if output_is_compressed else _lib.ENC_UNCOMPRESSED
    fl_in = (_lib.CHECK_INPUT if check_input_for_correctness else 0) | _lib.REJECT_INFINITY
    fl_out = (_lib.CHECK_INPUT if check_output_for_correctness else 0) | _lib.REJECT_INFINITY

    def chunk(amap, sec, name, start, count):
        off, total, size, group = sec[name]
        count = max(0, min(count, total - start))
        return group, np.asarray(amap[off + start * size: off + (start + count) * size])

    def check_before(name, start, count):
        # the reference always deserializes the challenge chunk (read_chunk, :398-410): even with CheckForCorrectness::No a bad
        # flag, a non-canonical coordinate or a point at infinity in `before` panics.  Validation only: nothing comes back.
        group, raw = chunk(input_map, sec_in, name, start, count)
        try:
            ctx.validate(group, raw, enc_in, fl_in)
        except _lib.P2BError as e:
            BatchedAccumulator._raise(e)

    def powers_ok(name, start, count, pair, wait=False, points=None):
        """same_ratio(power_pairs(after.<name>[start .. start + count)), pair): the chunk goes to the GPU in its file encoding
        (decompressed and checked there), the random coefficients are generated there, two points come back."""
        if points is None:
            group, raw = chunk(output_map, sec_out, name, start, count)
            enc, fl = enc_out, fl_out
        else:
            group, raw, enc, fl = 0, points, _lib.ENC_UNCOMPRESSED, 0
        if raw.size // _lib.enc_size(group, enc) < 2:
            return False                              # merge_pairs of nothing is (0, 0): same_ratio rejects zero
        try:
            s, sx = power_pairs(ctx, group, raw, None, rng, scalar_bits, enc, fl)
        except _lib.P2BError as e:
            BatchedAccumulator._raise(e)
        pending.append(pool.submit(same_ratio, (s, sx), pair) if group == 0 else pool.submit(same_ratio, pair, (s, sx)))
        done = [f for f in pending if f.done()]
        if wait:
            done = list(pending)
        return all(f.result() for f in done)          # a failed check stops the walk at the next chunk at the latest

    last_first = [None, None]
    for start in range(0, p.powers_length, p.batch_size):
        end = min(start + p.batch_size, p.powers_length) - 1
        if end == start:
            raise RuntimeError("Chunk does not have a min and max")            # the reference's panic (:462)
        size = end - start + 1 + (0 if end == p.powers_length - 1 else 1)      # one extra element: chunks overlap
        for name in ("tau_g1", "tau_g2", "alpha_g1", "beta_g1"):
            check_before(name, start, size)
        if not powers_ok("tau_g1", start, size, g2_pair):
            return False
        if not powers_ok("tau_g2", start, size, g1_pair):
            return False
        if not powers_ok("alpha_g1", start, size, g2_pair):
            return False
        if not powers_ok("beta_g1", start, size, g2_pair):
            return False
        if end == p.powers_length - 1:
            last_first[0] = pt(after("tau_g1", end, 1), 0, 64)
    for start in range(p.powers_length, p.powers_g1_length, p.batch_size):
        end = min(start + p.batch_size, p.powers_g1_length) - 1
        if end == start:
            raise RuntimeError("Chunk does not have a min and max")            # (:526)
        size = end - start + 1 + (0 if end == p.powers_g1_length - 1 else 1)
        check_before("tau_g1", start, size)
        if not powers_ok("tau_g1", start, size, g2_pair):
            return False
        if start == p.powers_length:
            last_first[1] = pt(after("tau_g1", start, 1), 0, 64)
    if last_first[0] is None or last_first[1] is None:
        return False
    return powers_ok(None, 0, 0, g2_pair, wait=True, points=np.frombuffer(last_first[0] + last_first[1], dtype=np.uint8))


BatchedAccumulator.verify_transformation = staticmethod(verify_transformation)


def _seed(rng):
    """32 bytes for the device-side coefficient generator (ChaCha20 keyed by them) from `rng` (system_rng by default)."""
    return bytes((rng or system_rng()).bytes(32))


def merge_pairs(ctx, group, v1, v2, scalars=None, rng=None, scalar_bits=253, in_enc=_lib.ENC_UNCOMPRESSED, flags=0):
    """(sum r_i v1_i, sum r_i v2_i): the random linear combination behind every same_ratio check of the verifier
    (powersoftau/src/utils.rs:112-130, phase2/src/utils.rs:59-105) in ONE pass on the GPU (p2b_g{1,2}_msm_pair: one sort of the
    shared coefficients, two bucket sets).  The reference draws the r_i from thread_rng; here they are either explicit
    (`scalars`, 32-byte big-endian each) or generated on the device from 32 bytes of `rng` (the OS CSPRNG by default).
    Returns two uncompressed points."""
    if scalars is not None:
        return ctx.msm_pair(group, v1, v2, scalars, in_enc=in_enc, flags=flags)
    _random_scalars(system_rng(), 0, scalar_bits)
    return ctx.msm_pair(group, v1, v2, None, _seed(rng), scalar_bits, in_enc, flags)


def power_pairs(ctx, group, v, scalars=None, rng=None, scalar_bits=253, in_enc=_lib.ENC_UNCOMPRESSED, flags=0):
    """merge_pairs(v[..n-1], v[1..]) (utils.rs:133-135): checks that consecutive elements have the same ratio.  One upload of
    v, one pass (p2b_g{1,2}_power_pairs)."""
    if scalars is not None:
        size = _lib.enc_size(group, in_enc)
        a = np.frombuffer(v, dtype=np.uint8) if not isinstance(v, np.ndarray) else v
        n = a.size // size
        sc = np.frombuffer(scalars, dtype=np.uint8) if not isinstance(scalars, np.ndarray) else scalars
        return ctx.power_pairs(group, a, sc[: (n - 1) * 32], in_enc=in_enc, flags=flags)
    _random_scalars(system_rng(), 0, scalar_bits)
    return ctx.power_pairs(group, v, None, _seed(rng), scalar_bits, in_enc, flags)


def prepare_phase2(ctx, accumulator_map, parameters, m, input_is_compressed=True, check_input_for_correctness=True,
                   g2_in_subgroup=False):
    """One iteration of powersoftau/src/bin/prepare_phase2.rs:62-241: the bytes of the file `phase1radix2m{m}` --
    alpha_g1, beta_g1, beta_g2, then the Lagrange coefficients (group iFFT of the first 2^m powers) in G1, G2, alpha*G1,
    beta*G1 and the H query tau^(i+d) G - tau^i G, all uncompressed.  `accumulator_map` is a response (compressed, the
    binary's input) or a challenge (uncompressed) including its 64-byte hash prefix."""
    try:
        return ctx.pot_prepare_phase2(accumulator_map, parameters.size, m, bool(input_is_compressed),
                                      bool(check_input_for_correctness), _lib.G2_SUBGROUP if g2_in_subgroup else 0)
    except _lib.P2BError as e:
        BatchedAccumulator._raise(e)
