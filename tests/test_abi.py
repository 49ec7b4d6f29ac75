"""CPU: the C-ABI library loads without a GPU, exports every symbol include/p2b.h declares, the ctypes binding covers
them all, and the product path fails loudly (no CPU fallback) when no CUDA device is present."""
import ctypes
import os
import re
import subprocess

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def header_symbols():
    src = open(os.path.join(ROOT, "include", "p2b.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(p2b_[a-z0-9_]+)\s*\(", src)))


def test_header_binding_and_exports_agree():
    from phase2_bn254_b200 import lib
    hdr = header_symbols()
    assert hdr == sorted(lib.SYMBOLS)
    so = lib.load()
    for name in hdr:
        assert hasattr(so, name), name
    nm = subprocess.check_output(["nm", "-D", "--defined-only", lib.LIB_PATH]).decode()
    exported = sorted(set(re.findall(r" T (p2b_[a-z0-9_]+)", nm)))
    assert exported == hdr, "exported symbols differ from include/p2b.h"
    assert b"sm_100a" in so.p2b_version()


def test_no_torch_types_in_the_abi():
    src = open(os.path.join(ROOT, "include", "p2b.h")).read()
    assert "torch" not in src and "at::" not in src and "#include <cuda" not in src


def test_library_is_built_for_sm_100a():
    from phase2_bn254_b200 import lib
    out = subprocess.run(["cuobjdump", "-lelf", lib.LIB_PATH], capture_output=True, text=True)
    if out.returncode != 0:
        pytest.skip("cuobjdump unavailable")
    assert "sm_100a" in out.stdout


def test_fails_loudly_without_gpu():
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    from phase2_bn254_b200 import lib
    with pytest.raises(lib.P2BError) as e:
        lib.Context(0)
    assert e.value.code == lib.ECUDA
    so = lib.load()
    so.p2b_g1_msm.restype = ctypes.c_int
    assert so.p2b_g1_msm(None, None, None, 0, None) == lib.EARG           # null ctx is an argument error, not a crash


def test_geometry_matches_reference_table():
    """powersoftau/src/parameters.rs:72-107 at the configs of SURVEY.md section 8."""
    from phase2_bn254_b200 import lib
    from phase2_bn254_b200.powersoftau import CeremonyParams
    so = lib.load()
    for size, acc, contrib in ((10, 393344, 197472), (20, 402653312, 201327456), (26, 25769803904, 12884902752),
                               (28, 103079215232, 51539608416)):
        p = CeremonyParams(size, 256)
        assert p.accumulator_size == acc and p.contribution_size == contrib
        assert so.p2b_pot_accumulator_size(size, 0) == acc
        assert so.p2b_pot_accumulator_size(size, 1) == contrib - 768


def test_generate_initial_matches_oracle(oracle):
    """Host-only mirror of BatchedAccumulator::generate_initial (batched_accumulator.rs:1295-1347), both encodings."""
    import hashlib
    import numpy as np
    from phase2_bn254_b200.powersoftau import BatchedAccumulator, CeremonyParams
    for size in (1, 4, 7):
        p = CeremonyParams(size, 16)
        for compressed in (False, True):
            n = p.contribution_size - p.public_key_size if compressed else p.accumulator_size
            out = np.zeros(n, dtype=np.uint8)
            out[:64] = np.frombuffer(hashlib.blake2b(b"").digest(), dtype=np.uint8)
            BatchedAccumulator.generate_initial(out, compressed, p)
            assert out.tobytes() == oracle.pot_generate_initial(size, compressed)


def test_params_layout_host_logic():
    import struct
    from phase2_bn254_b200.phase2 import params_layout
    body = bytes(64 + 64 + 128 + 128 + 64 + 128)
    vecs = b"".join(struct.pack(">I", n) + bytes(n * sz) for n, sz in ((2, 64), (3, 64), (4, 64), (1, 64), (1, 64), (1, 128)))
    buf = body + vecs + bytes(64) + struct.pack(">I", 1) + bytes(384)
    lay = params_layout(buf)
    assert lay["h"][1] == 3 and lay["l"][1] == 4 and lay["contributions"][1] == 1 and lay["b_g2"][2:] == (128, 1)
    with pytest.raises(ValueError):
        params_layout(buf + b"x")
    with pytest.raises(ValueError):
        params_layout(buf[:-1])


def test_size_helpers_and_null_ctx():
    from phase2_bn254_b200 import lib
    so = lib.load()
    for m in (0, 1, 10, 20):
        assert so.p2b_pot_radix_file_size(m) == 192 + 384 * (1 << m)          # prepare_phase2.rs:158-240 layout
    import ctypes
    for name in ("p2b_g1_group_fft", "p2b_pot_prepare_phase2", "p2b_pot_decompress", "p2b_g2_recode", "p2b_fr_fft", "p2b_sync"):
        fn = getattr(so, name)
        fn.restype = ctypes.c_int
        nargs = len(fn.argtypes)
        assert fn(*([None] + [0] * (nargs - 1))) == lib.EARG, name


def test_flag_and_code_constants_match_the_header():
    """The binding's flag bits, encodings and error codes are the header's enum values (include/p2b.h)."""
    import re
    from phase2_bn254_b200 import lib
    text = open(os.path.join(ROOT, "include", "p2b.h")).read()
    vals = {m.group(1): int(m.group(2)) for m in re.finditer(r"\b(P2B_[A-Z0-9_]+)\s*=\s*(\d+)", text)}
    for name in ("CHECK_INPUT", "REJECT_INFINITY", "G2_SUBGROUP", "G2_EXACT", "ENC_UNCOMPRESSED", "ENC_COMPRESSED", "ENC_RAW_MONT_LE",
                 "EARG", "EDECODE", "EINFINITY_IN", "EINFINITY_OUT", "DEC_NOT_ON_CURVE", "DEC_COORDINATE",
                 "DEC_UNEXPECTED_INFORMATION", "DEC_UNEXPECTED_COMPRESSION_MODE"):
        assert getattr(lib, name) == vals["P2B_" + name], name
    assert vals["P2B_G2_EXACT"] == 8 and vals["P2B_G2_SUBGROUP"] == 4
