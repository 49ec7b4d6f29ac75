"""Host-side mirror of the bellman entry points on the hot path, backed by libp2b.so.

  multiexp / dense_multiexp      bellman/src/multiexp.rs:330-475   -> multiexp()
  EvaluationDomain               bellman/src/domain.rs:52-205      -> EvaluationDomain
Scalars are Python ints in [0, r); points are uncompressed wire bytes.
"""
import numpy as np

from . import lib as _lib

FR_MODULUS = 21888242871839275222246405745257275088548364400416034343698204186575808495617
FR_S = 28


def multiexp(ctx, group, bases, exponents):
    """sum exponents[i] * bases[i] -> uncompressed wire bytes of the (affine) result."""
    sc = b"".join(int(e).to_bytes(32, "big") for e in exponents) if not isinstance(exponents, (bytes, np.ndarray)) \
        else exponents
    return ctx.msm(group, bases, sc)


class EvaluationDomain:
    """Radix-2 domain over Fr; coeffs are padded with zeros to the next power of two (domain.rs:52-99)."""

    def __init__(self, ctx, coeffs):
        n = len(coeffs)
        if n > (1 << FR_S) - 1:
            raise ValueError("PolynomialDegreeTooLarge")
        m, exp = 1, 0
        while m < n:
            m *= 2
            exp += 1
        self.ctx, self.exp = ctx, exp
        buf = bytearray(32 * m)
        for i, c in enumerate(coeffs):
            buf[32 * i: 32 * i + 32] = int(c).to_bytes(32, "big")
        self.data = np.frombuffer(bytes(buf), dtype=np.uint8)

    @classmethod
    def from_coeffs(cls, ctx, coeffs):
        return cls(ctx, coeffs)

    def into_coeffs(self):
        b = self.data.tobytes()
        return [int.from_bytes(b[i: i + 32], "big") for i in range(0, len(b), 32)]

    def fft(self):
        self.data = self.ctx.fr_fft(self.data, False, False)

    def ifft(self):
        self.data = self.ctx.fr_fft(self.data, True, False)

    def coset_fft(self):
        self.data = self.ctx.fr_fft(self.data, False, True)

    def icoset_fft(self):
        self.data = self.ctx.fr_fft(self.data, True, True)
