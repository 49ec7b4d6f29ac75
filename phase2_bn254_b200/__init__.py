"""phase2_bn254_b200 -- B200 (sm_100a) compute core for the phase2-bn254 contribution hot path.

Layout
  csrc/           hand-written CUDA kernels + the C-ABI shared library (include/p2b.h -> libp2b.so)
  lib.py          ctypes binding of the C ABI (no CPU fallback)
  powersoftau.py  host-side mirror of powersoftau's BatchedAccumulator::transform interface
  phase2.py       host-side mirror of phase2's MPCParameters::{read, write, contribute}
  bellman.py      host-side mirror of bellman's multiexp / EvaluationDomain entry points
"""
from . import lib  # noqa: F401
from .lib import Context, P2BError  # noqa: F401
