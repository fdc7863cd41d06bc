"""bench.py's contract on the CPU side: the reference arm (the reference's algorithm on the host cores) prints exactly
one JSON line with the keys the driver reads, and ranks other than 0 of a multi-rank launch exit 0 without work."""
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _run(env_extra, *args):
    env = dict(os.environ, **env_extra)
    return subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", *args], env=env, cwd=ROOT,
                          capture_output=True, text=True, timeout=600)


def test_reference_arm_other_ranks_exit_without_output():
    out = _run({"RANK": "1", "WORLD_SIZE": "2", "LOCAL_RANK": "1"}, "--gpus", "2", "--steps", "1", "--warmup", "0")
    assert out.returncode == 0 and out.stdout.strip() == ""


def test_reference_arm_prints_one_json_line():
    # a small sample: 256 columns of the 400-row workload, one step
    out = _run({"RANK": "0", "WORLD_SIZE": "1"}, "--steps", "1", "--warmup", "0", "--cpu-sample-columns", "256", "--cpu-procs", "2")
    assert out.returncode == 0, out.stderr[-2000:]
    lines = [l for l in out.stdout.splitlines() if l.strip()]
    assert len(lines) == 1
    line = json.loads(lines[0])
    assert line["impl"] == "reference" and line["unit"] == "columns/s" and line["higher_is_better"] is True
    assert line["gpu_launches"] == 0 and line["value"] > 0 and line["steps"] == 1
    assert line["cpu_baseline"]["kind"] in ("reference", "port") and line["cpu_baseline"]["cores"] >= 1
    staged = os.path.exists(os.path.join(ROOT, "oracle", "_ref", "motes", "tracing.code")) or os.path.isdir("/root/reference/src/motes")
    assert line["cpu_baseline"]["kind"] == ("reference" if staged else "port")
    assert line["e2e"] == {"value": line["value"], "unit": "columns/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}
    assert "workload" in line["config"] and line["dtype"] == "f64" and line["data"] == "synthetic"
    assert line["config"]["workload"].startswith("C2 FORS2-like 4096x400") and line["config"]["frames_per_gpu"] == 128


def test_cpu_legs_do_not_load_the_cuda_library():
    """The CPU arm imports the frame generator and the checker only: libmotes_b200.so stays out of the process."""
    code = ("import sys; sys.path.insert(0, %r); import bench; "
            "print(bench._cpu_worker((1, 96, 256, 0, 0.0, 'on', 0))[0]); "
            "import motes_b200; print('lib' if 'motes_b200._lib' in sys.modules else 'nolib')") % ROOT
    out = subprocess.run([sys.executable, "-c", code], capture_output=True, text=True, cwd=ROOT, timeout=300)
    assert out.returncode == 0, out.stderr[-2000:]
    assert out.stdout.split() == ["256", "nolib"]
