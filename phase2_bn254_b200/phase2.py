"""Host-side mirror of the phase2 crate's MPCParameters contribution interface, backed by libp2b.so.

Mirrors  MPCParameters::{read, write, contribute}  phase2/src/parameters.rs:414-522,663-703 and the
PublicKey wire form phase2/src/keypair.rs:50-105.  Parameters are kept in their serialized form
(bellman/src/groth16/mod.rs:252-285): the GPU path consumes and produces wire bytes directly.
"""
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


class MPCParameters:
    def __init__(self, data):
        self.data = np.array(np.frombuffer(bytes(data), dtype=np.uint8)) if not isinstance(data, np.ndarray) else data

    @classmethod
    def read(cls, reader, disallow_points_at_infinity=False, checked=True, ctx=None):
        """MPCParameters::read (phase2/src/parameters.rs:682-703 -> Parameters::read, groth16/mod.rs:287-383): every point
        is range-checked, `checked` adds is_on_curve, `disallow_points_at_infinity` rejects infinity -- done section by
        section with the GPU bulk codec.  Raises IOError("InvalidData ...") like the reference's io::Error."""
        buf = reader if isinstance(reader, (bytes, bytearray, memoryview, np.ndarray)) else reader.read()
        self = cls(buf)
        ctx = ctx or _lib.Context(0)
        flags = (_lib.CHECK_INPUT if checked else 0) | (_lib.REJECT_INFINITY if disallow_points_at_infinity else 0)
        lay = params_layout(self.data)
        for name, (off, n, size, group) in lay.items():
            if group is None or n == 0:
                continue
            try:
                ctx.recode(group, self.data[off: off + n * size], _lib.ENC_UNCOMPRESSED, _lib.ENC_UNCOMPRESSED, flags)
            except _lib.P2BError as e:
                if e.code in (_lib.EDECODE, _lib.EINFINITY_IN):
                    what = "point at infinity" if e.code == _lib.EINFINITY_IN else \
                        {1: "NotOnCurve", 2: "CoordinateDecodingError", 3: "UnexpectedInformation",
                         4: "UnexpectedCompressionMode"}.get(e.sub, "decoding error")
                    raise IOError("InvalidData: %s in %s[%d]" % (what, name, e.index))
                raise
        return self

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
