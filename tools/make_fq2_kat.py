"""Transcribes the reference's hard-coded Fq2 tables (pairing/src/bn256/fq.rs:106-431) into tests/golden/fq2_frobenius_kat.json.
Run in the build container (reads /root/reference); the tests only read the committed JSON."""
import json
import os
import re

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
src = open("/root/reference/pairing/src/bn256/fq.rs").read().split("\n")


def grab(name):
    i = [k for k, l in enumerate(src) if name in l and "pub const" in l][0]
    vals, j = [], i
    while True:
        j += 1
        vals += [int(v, 16) for v in re.findall(r"0x[0-9a-f]+", src[j])]
        if src[j].startswith("];") or src[j].startswith("};"):
            break
    return i + 1, j + 1, vals


out = {}
for nm in ("XI_TO_Q_MINUS_1_OVER_2", "FROBENIUS_COEFF_FQ6_C1", "FROBENIUS_COEFF_FQ6_C2", "FROBENIUS_COEFF_FQ12_C1"):
    a, b, v = grab(nm)
    out[nm] = {"lines": "pairing/src/bn256/fq.rs:%d-%d" % (a, b), "limbs": [hex(x) for x in v]}
json.dump({"_what": "The reference's hard-coded Fq2 tables (Montgomery limbs, c0 then c1 per entry), transcribed by "
                    "tools/make_fq2_kat.py from the lines given; every entry is a power of xi = 9 + u, so they are known-answer "
                    "tests for any Fq / Fq2 implementation", "tables": out},
          open(os.path.join(ROOT, "tests", "golden", "fq2_frobenius_kat.json"), "w"), indent=1)
