"""bench.py contract checks that need no GPU: the reference arm (the CPU restatement timed on the host cores)
prints one JSON line with the keys the driver reads; under a multi-rank launch only rank 0 prints; the B200 arm
refuses to run without a device instead of falling back to the CPU."""
import json
import os
import subprocess
import sys
from pathlib import Path

ROOT = Path(__file__).resolve().parent.parent


def run(args, env=None):
    e = dict(os.environ)
    e.update(env or {})
    return subprocess.run([sys.executable, str(ROOT / "bench.py"), *args], capture_output=True, text=True, env=e,
                          timeout=600)


def test_reference_arm_line():
    r = run(["--impl", "reference", "--workload", "c1", "--steps", "2", "--warmup", "1"])
    assert r.returncode == 0, r.stderr[-2000:]
    line = json.loads(r.stdout.strip().splitlines()[-1])
    for key in ("impl", "metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better",
                "scaling", "vs_baseline", "dtype", "data", "config", "cpu_baseline", "e2e"):
        assert key in line, key
    assert line["impl"] == "reference" and line["value"] > 0 and line["unit"] == "GB/s"
    assert line["cpu_baseline"]["kind"] == "port" and line["cpu_baseline"]["cores"] >= 1
    assert line["e2e"]["h2d_bytes_per_step"] == 0 and line["e2e"]["d2h_bytes_per_step"] == 0
    assert "workload" in line["config"]


def test_reference_arm_only_rank0_prints():
    r = run(["--impl", "reference", "--workload", "c1", "--steps", "1", "--warmup", "1", "--gpus", "2"],
            env={"RANK": "1", "WORLD_SIZE": "2", "LOCAL_RANK": "1"})
    assert r.returncode == 0 and r.stdout.strip() == ""


def test_b200_arm_needs_a_device():
    import torch
    if torch.cuda.is_available():
        return
    r = run(["--workload", "c1", "--steps", "1"])
    assert r.returncode != 0 and "no CPU fallback" in (r.stderr + r.stdout)
