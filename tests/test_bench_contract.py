"""bench.py's CPU arm on a tiny size: the JSON line carries the keys the driver reads (same metric / unit / config shape as the GPU
arm, `impl: reference`, `cpu_baseline`, a self-describing `e2e`), and the window rule is bellman's."""
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_reference_arm_json_contract():
    out = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--log-n", "12", "--steps", "2",
                          "--warmup", "1"], capture_output=True, text=True, timeout=300)
    assert out.returncode == 0, out.stderr
    lines = [l for l in out.stdout.splitlines() if l.strip()]
    assert len(lines) == 1
    d = json.loads(lines[0])
    assert d["impl"] == "reference" and d["metric"] == "BN254 G1 MSM throughput" and d["unit"] == "Mscalar-mul/s"
    assert d["higher_is_better"] is True and d["vs_baseline"] is None and d["scaling"] == "weak"
    assert d["config"]["workload"] == "g1_msm_2^12" and d["config"]["terms_per_gpu"] == 4096
    assert d["config"]["window_bits"] == 9 and d["config"]["windows"] == 29           # ceil(ln 4096) = 9, ceil(254 / 9) = 29
    assert d["steps"] == 2 and d["value"] > 0 and d["ms_per_step"] > 0
    assert d["cpu_baseline"]["kind"] == "port" and d["cpu_baseline"]["cores"] >= 1 and d["cpu_baseline"]["value"] == d["value"]
    assert d["e2e"] == {"value": d["value"], "unit": d["unit"], "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}


def test_reference_arm_other_ranks_exit_quietly():
    env = dict(os.environ, RANK="1", WORLD_SIZE="2", LOCAL_RANK="1")
    out = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--gpus", "2", "--log-n", "12"],
                         capture_output=True, text=True, timeout=120, env=env)
    assert out.returncode == 0 and out.stdout.strip() == ""


def test_window_rules():
    sys.path.insert(0, ROOT)
    import bench
    assert bench.bellman_windows(1 << 20) == (14, 19) and bench.bellman_windows(1 << 26) == (19, 14) and bench.bellman_windows(16) == (3, 85)
    assert bench.msm_windows(26) == 14 and bench.msm_windows(20) == 17           # csrc/msm_impl.cuh msm_geometry
