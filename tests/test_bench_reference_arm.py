"""CPU: the `bench.py --impl reference` arm keeps the bench contract — exactly K timed steps after W warm-up steps,
the same metric / unit / config as the GPU arm, zero copies, rank 0 alone prints (the other ranks exit 0 silently)."""
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _run(extra_env=None, args=("--steps", "3", "--warmup", "1")):
    env = dict(os.environ)
    env.update(extra_env or {})
    p = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", *args],
                       stdout=subprocess.PIPE, stderr=subprocess.PIPE, text=True, timeout=300, env=env, cwd=ROOT)
    assert p.returncode == 0, p.stderr[-2000:]
    return [l for l in p.stdout.splitlines() if l.startswith("{")]


def test_reference_arm_line_honours_steps_and_warmup():
    lines = _run()
    assert len(lines) == 1
    d = json.loads(lines[0])
    assert d["impl"] == "reference"
    assert (d["steps"], d["warmup"]) == (3, 1)
    assert d["metric"] == "agent-steps/sec" and d["unit"] == "agent-steps/s" and d["higher_is_better"] is True
    assert d["config"]["name"] == "cleanup8" and "cleanup_new n=8" in d["config"]["workload"]
    assert d["gpu_launches"] == 0
    assert d["e2e"] == {"value": d["value"], "unit": d["unit"], "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}
    cb = d["cpu_baseline"]
    assert cb["kind"] in ("reference", "port") and cb["value"] == d["value"] and cb["cores"] >= 1 and cb["sample"]
    assert d["value"] > 0 and d["ms_per_step"] > 0
    if cb["kind"] == "reference":
        # one step of this arm = R env steps in every one-env process: value = P * n * R / ms_per_step
        R = int(d["step_definition"].split()[0])
        assert abs(d["value"] - cb["cores"] * 8 * R / (d["ms_per_step"] * 1e-3)) <= 1e-6 * d["value"]
        assert "x %d steps" % (3 * R) in cb["sample"]


def test_reference_arm_other_ranks_print_nothing():
    assert _run({"RANK": "1", "LOCAL_RANK": "1", "WORLD_SIZE": "2"}, ("--gpus", "2", "--steps", "3", "--warmup", "1")) == []
