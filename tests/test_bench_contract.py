"""CPU: the parts of the bench.py contract that do not need a GPU - the reference arm prints exactly ONE JSON line on stdout
with the contract's keys (everything else goes to stderr), and the workload table names BASELINE.json's configurations."""
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_reference_arm_prints_one_json_line():
    r = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--workload", "c1", "--steps", "1", "--warmup", "1"],
                       capture_output=True, text=True, timeout=600, cwd=ROOT)
    assert r.returncode == 0, r.stderr[-2000:]
    lines = [l for l in r.stdout.split("\n") if l.strip()]
    assert len(lines) == 1, r.stdout
    d = json.loads(lines[0])
    assert d["impl"] == "reference" and d["metric"] == "train images/sec" and d["unit"] == "images/s" and d["higher_is_better"] is True
    assert d["value"] > 0 and d["n_gpus"] == 1 and d["steps"] == 1
    assert d["cpu_baseline"]["kind"] == "port" and d["cpu_baseline"]["cores"] >= 1 and d["cpu_baseline"]["value"] == d["value"]
    assert d["e2e"] == {"value": d["value"], "unit": "images/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}
    assert "SVHN-shape 32x32x3" in d["config"]["workload"]


def test_reference_arm_other_ranks_exit_quietly():
    env = dict(os.environ, RANK="1", WORLD_SIZE="2", LOCAL_RANK="1")
    r = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--gpus", "2", "--steps", "1"],
                       capture_output=True, text=True, timeout=120, cwd=ROOT, env=env)
    assert r.returncode == 0 and r.stdout.strip() == ""


def test_workloads_name_the_baseline_configs():
    sys.path.insert(0, ROOT)
    import bench
    with open(os.path.join(ROOT, "BASELINE.json")) as f:
        B = json.load(f)
    assert bench.WORKLOADS["c2"][:3] == ("lgvae", 64, 256) and "--beta 120 --patch_size 8 -no_label" in bench.WORKLOADS["c2"][6]
    assert "--beta 120 --patch_size 8 -no_label" in B["configs"][1]
    assert bench.WORKLOADS["c1"][:3] == ("lgvae", 32, 64) and bench.WORKLOADS["c3"][0] == "lggmvae" and bench.WORKLOADS["c4"][:2] == ("lggmvae", 64)
