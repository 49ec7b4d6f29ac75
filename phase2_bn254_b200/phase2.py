"""Host-side mirror of the phase2 crate's MPCParameters contribution interface, backed by libp2b.so.

Mirrors  MPCParameters::{read, write, contribute}  phase2/src/parameters.rs:414-522,663-703 and the
PublicKey wire form phase2/src/keypair.rs:50-105.  Parameters are kept in their serialized form
(bellman/src/groth16/mod.rs:252-285): the GPU path consumes and produces wire bytes directly.
"""
import numpy as np

from . import lib as _lib


class MPCParameters:
    def __init__(self, data):
        self.data = np.array(np.frombuffer(bytes(data), dtype=np.uint8)) if not isinstance(data, np.ndarray) else data

    @classmethod
    def read(cls, reader, disallow_points_at_infinity=False, checked=True):
        """`reader` is a file object or bytes.  Point validation of the untouched sections is the caller's
        (reference: Parameters::read, groth16/mod.rs:287-383); h, l, delta are validated on the GPU in contribute."""
        buf = reader if isinstance(reader, (bytes, bytearray, memoryview, np.ndarray)) else reader.read()
        return cls(buf)

    def write(self, writer):
        writer.write(self.data.tobytes())

    def contribute(self, delta, s_g1, r_g2=None, ctx=None, hash_to_g2=None):
        """Contributes `delta` (int in [1, r)).  The reference draws delta, s = G1::rand and r = hash_to_g2(transcript)
        from its ChaCha RNG (parameters.rs:860-908); here they are explicit: pass r_g2, or a `hash_to_g2` callable
        mapping the 64-byte transcript to a 128-byte uncompressed G2 point.  Returns the 64-byte contribution hash."""
        ctx = ctx or _lib.Context(0)
        d = np.frombuffer(int(delta).to_bytes(32, "big"), dtype=np.uint8)
        if r_g2 is None:
            if hash_to_g2 is None:
                raise ValueError("need r_g2 or hash_to_g2")
            r_g2 = hash_to_g2(ctx.phase2_transcript(self.data, d, s_g1))
        out, h = ctx.phase2_contribute(self.data, d, s_g1, r_g2)
        self.data = out
        return h
