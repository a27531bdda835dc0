"""bench.py keeps the driver's contract: the reference arm runs on CPU and prints ONE JSON line with the agreed keys."""
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_reference_arm_json_line():
    out = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--steps", "2", "--warmup", "1"],
                         check=True, capture_output=True, text=True, cwd=ROOT, timeout=600).stdout.strip().splitlines()
    assert len(out) == 1
    d = json.loads(out[0])
    assert d["impl"] == "reference" and d["metric"] == "4D TV-FISTA Gvoxel-iter/s" and d["unit"] == "Gvoxel*iter/s"
    assert d["higher_is_better"] is True and d["value"] > 0 and d["steps"] == 2 and d["gpu_launches"] == 0
    assert d["cpu_baseline"]["kind"] in ("reference", "port") and d["cpu_baseline"]["cores"] >= 1
    assert d["cpu_baseline"]["value"] == d["value"] == d["e2e"]["value"]
    assert d["e2e"]["h2d_bytes_per_step"] == 0 and d["e2e"]["d2h_bytes_per_step"] == 0
    assert "workload" in d["config"] and d["vs_baseline"] is None and d["dtype"] == "f32"
    assert d["warmup"] == 1 and d["config"]["cpu_sample_shape"] in ([256, 256, 128, 128], [32, 32, 128, 128])
    assert "x".join(map(str, d["config"]["cpu_sample_shape"])) in d["cpu_baseline"]["sample"]


def test_reference_arm_uses_all_cores_under_torchrun():
    """torchrun exports OMP_NUM_THREADS=1 to every rank; rank 0's reference arm must override it (round 1 timed the
    N > 1 reference arm on ONE thread, which inflated the driver's ratio 8.5x)."""
    env = dict(os.environ, RANK="0", WORLD_SIZE="2", OMP_NUM_THREADS="1", CYTVDN_BENCH_CPU_FULL="0")
    out = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--gpus", "2", "--steps", "2",
                          "--warmup", "1"], check=True, capture_output=True, text=True, cwd=ROOT, env=env, timeout=300).stdout
    d = json.loads(out.strip().splitlines()[-1])
    assert d["cpu_baseline"]["cores"] == (os.cpu_count() or 1) and d["n_gpus"] == 2
    assert d["config"]["cpu_sample_shape"] == [32, 32, 128, 128] and "cannot hold" in d["config"]["cpu_sample"]


def test_reference_arm_other_ranks_stay_silent():
    env = dict(os.environ, RANK="1", WORLD_SIZE="2")
    r = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--gpus", "2", "--steps", "1",
                        "--warmup", "1"], capture_output=True, text=True, cwd=ROOT, env=env, timeout=300)
    assert r.returncode == 0 and r.stdout.strip() == ""
