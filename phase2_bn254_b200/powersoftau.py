"""Host-side mirror of the powersoftau crate's contribution interface, backed by libp2b.so.

Mirrors (names, argument meaning, error behaviour):
  CeremonyParams         powersoftau/src/parameters.rs:38-120
  UseCompression, CheckForCorrectness   parameters.rs:127-140
  DeserializationError   parameters.rs:143-170
  PrivateKey             powersoftau/src/keypair.rs:47-51
  BatchedAccumulator.transform   powersoftau/src/batched_accumulator.rs:1119-1292
The maps are numpy uint8 arrays (np.memmap works), exactly the byte layout of the challenge / response files.
"""
import hashlib
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


def merge_pairs(ctx, group, v1, v2, scalars):
    """(sum r_i v1_i, sum r_i v2_i): the random linear combination behind every same_ratio check of the verifier
    (powersoftau/src/utils.rs:112-130, phase2/src/utils.rs:59-105); two Pippenger MSMs on the GPU.  The reference draws
    the r_i from thread_rng; here they are explicit (32-byte big-endian each).  Returns two uncompressed points."""
    return ctx.msm(group, v1, scalars), ctx.msm(group, v2, scalars)


def power_pairs(ctx, group, v, scalars):
    """merge_pairs(v[..n-1], v[1..]) (utils.rs:133-135): checks that consecutive elements have the same ratio."""
    size = 128 if group else 64
    a = np.frombuffer(v, dtype=np.uint8) if not isinstance(v, np.ndarray) else v
    n = a.size // size
    sc = np.frombuffer(scalars, dtype=np.uint8) if not isinstance(scalars, np.ndarray) else scalars
    return merge_pairs(ctx, group, a[: (n - 1) * size], a[size:], sc[: (n - 1) * 32])


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
