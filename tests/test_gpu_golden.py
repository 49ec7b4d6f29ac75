"""GPU parity against the committed golden vectors (tests/golden/vectors.json; python big-int arithmetic)."""
import hashlib
import json
import os

import numpy as np
import pytest

pytestmark = pytest.mark.gpu
GOLD = json.load(open(os.path.join(os.path.dirname(__file__), "golden", "vectors.json")))
hx = bytes.fromhex


@pytest.mark.parametrize("group", [0, 1])
def test_batch_mul_golden(ctx, group):
    g = GOLD["g%d_batch_mul" % (group + 1)]
    pts, sc = hx(g["points"]), hx(g["scalars"])
    assert ctx.batch_mul(group, pts, sc).tobytes().hex() == g["out"]
    assert ctx.batch_mul(group, pts, sc, 0, 1).tobytes().hex() == g["out_compressed"]
    assert ctx.batch_mul(group, hx(g["points_compressed"]), sc, 1, 0).tobytes().hex() == g["out"]
    b = GOLD["g%d_broadcast" % (group + 1)]
    assert ctx.batch_mul(group, pts, hx(b["scalar"])).tobytes().hex() == b["out"]
    assert ctx.msm(group, pts, sc).hex() == GOLD["g%d_msm" % (group + 1)]["out"]


def test_pot_transform_golden(ctx, oracle):
    from phase2_bn254_b200.powersoftau import BatchedAccumulator, CeremonyParams, PrivateKey
    p = GOLD["pot"]
    params = CeremonyParams(p["size"], p["batch"])
    ch = np.frombuffer(oracle.pot_generate_initial(p["size"]), dtype=np.uint8)
    assert hashlib.blake2b(ch.tobytes()).hexdigest() == p["challenge_blake2b"]
    key = PrivateKey(*[int(x, 16) for x in p["keys"]])
    resp = np.zeros(params.contribution_size, dtype=np.uint8)
    BatchedAccumulator.transform(ch, resp, False, True, False, key, params, ctx=ctx)
    end = params.contribution_size - params.public_key_size
    assert resp[64:end].tobytes().hex() == p["response_body"]
    nxt = np.zeros(params.accumulator_size, dtype=np.uint8)
    BatchedAccumulator.transform(ch, nxt, False, False, False, key, params, ctx=ctx)
    assert nxt[64:].tobytes().hex() == p["challenge2_body"]
    key2 = PrivateKey(*[int(x, 16) for x in p["keys2"]])
    resp2 = np.zeros(params.contribution_size, dtype=np.uint8)
    BatchedAccumulator.transform(nxt, resp2, False, True, True, key2, params, ctx=ctx)
    assert resp2[64:end].tobytes().hex() == p["response2_body"]
    resp3 = np.zeros(params.contribution_size, dtype=np.uint8)
    BatchedAccumulator.transform(resp[:end], resp3, True, True, True, key2, params, ctx=ctx)
    assert resp3[64:end].tobytes().hex() == p["response2_body"]


def test_phase2_contribute_golden(ctx):
    from phase2_bn254_b200.phase2 import MPCParameters
    p = GOLD["phase2"]
    mp = MPCParameters.read(hx(p["params"]))
    c = p["c1"]
    h = mp.contribute(int(c["delta"], 16), hx(c["s"]), hx(c["r"]), ctx=ctx)
    assert mp.data.tobytes().hex() == c["out"] and h.hex() == c["hash"]
    c = p["c2"]
    h = mp.contribute(int(c["delta"], 16), hx(c["s"]), hx(c["r"]), ctx=ctx)
    assert hashlib.blake2b(mp.data.tobytes()).hexdigest() == c["out_blake2b"] and h.hex() == c["hash"]
