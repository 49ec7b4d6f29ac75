"""Rebuilds profiles/traffic.json from an ncu metrics CSV of the bucket accumulation of ONE whole MSM
(ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,sm__pipe_fmaheavy_cycles_active... -k
regex:k_msm_accumulate, see tools/gpu_r2b.sh) and stamps it with the hashes of the kernel sources it was captured from, so
that bench.py can refuse to print a stale figure next to a live one.

    python tools/update_traffic.py profiles/r2a/ncu_accumulate_2p26_metrics.csv 26
"""
import csv
import hashlib
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
SOURCES = ["phase2_bn254_b200/csrc/msm_impl.cuh", "phase2_bn254_b200/csrc/xyzz.cuh", "phase2_bn254_b200/csrc/fp.cuh"]


def source_hash():
    h = hashlib.sha256()
    for f in SOURCES:
        h.update(open(os.path.join(ROOT, f), "rb").read())
    return h.hexdigest()[:16]


def main():
    path, log_n = sys.argv[1], int(sys.argv[2])
    rows, hdr, per = list(csv.reader(open(path))), None, {}
    for r in rows:
        if r and r[0] == "ID":
            hdr = r
            continue
        if hdr and len(r) == len(hdr):
            d = dict(zip(hdr, r))
            per.setdefault(d["ID"], {})[d["Metric Name"]] = float(d["Metric Value"].replace(",", ""))
    rd = sum(v.get("dram__bytes_read.sum", 0) for v in per.values())
    wr = sum(v.get("dram__bytes_write.sum", 0) for v in per.values())
    t = sum(v.get("gpu__time_duration.sum", 0) for v in per.values())
    busy = sum(v.get("sm__pipe_fmaheavy_cycles_active.avg.pct_of_peak_sustained_elapsed", 0) * v.get("gpu__time_duration.sum", 0)
               for v in per.values()) / t
    p = os.path.join(ROOT, "profiles", "traffic.json")
    try:
        out = json.load(open(p))
    except (OSError, ValueError):
        out = {}
    key = "k_msm_accumulate<Fq>@2^%d" % log_n
    out[key] = int(rd + wr)
    out[key + ":fmaheavy_busy"] = round(busy / 100, 4)
    out[key + ":launches"] = len(per)
    out[key + ":kernel_ms_under_ncu"] = round(t / 1e6, 3)
    out[key + ":capture"] = {"file": os.path.relpath(os.path.abspath(path), ROOT), "source_sha256_16": source_hash(), "sources": SOURCES,
                             "commit": subprocess.run(["git", "-C", ROOT, "rev-parse", "--short", "HEAD"], capture_output=True,
                                                      text=True).stdout.strip(),
                             "what": "sum over ALL accumulate launches of one MSM (not an extrapolation): dram__bytes_read.sum + "
                                     "dram__bytes_write.sum; busy = time-weighted sm__pipe_fmaheavy_cycles_active"}
    json.dump(out, open(p, "w"), indent=1)
    print(key, out[key], out[key + ":fmaheavy_busy"])


if __name__ == "__main__":
    main()
