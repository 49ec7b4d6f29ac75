import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "oracle"))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
import oracle as oc
from util import *
from phase2_bn254_b200 import lib
ctx = lib.Context(0)
P = random_points(oc, 0, 3, 1)
for k in [1, 65]:
    got = ctx.msm(0, P[:64], be(k)); exp = oc.msm(0, P[:64], be(k))
    print("k=%x" % k, got == exp, got.hex()[:16], exp.hex()[:16], P[:8].hex())
