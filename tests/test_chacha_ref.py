"""The python ChaCha20 used to check the library's device-generated verifier coefficients, pinned to RFC 7539 section 2.3.2."""
from util import chacha20_block, device_scalars


def test_rfc7539_block_vector():
    key = [int.from_bytes(bytes(range(4 * i, 4 * i + 4)), "little") for i in range(8)]
    out = chacha20_block(key, 1, 0x09000000, 0x4a000000, 0)
    assert out.hex() == ("10f1e7e4d13b5915500fdd1fa32071c4c7d1f4c733c068030422aa9ac3d46c4e"
                         "d2826446079faa0914c2d705d98b02a2b5129cd1de164eb9cbd083e8a2503c4e")


def test_device_scalar_layout():
    seed = bytes(range(32))
    a = device_scalars(seed, 5, 253)
    assert len(a) == 160 and all(a[32 * i] < 0x20 for i in range(5))
    b = device_scalars(seed, 5, 128)
    assert all(not any(b[32 * i: 32 * i + 16]) for i in range(5)) and b[16:32] == a[16:32]
    assert device_scalars(seed, 2, 253, first=3) == a[96:160]
