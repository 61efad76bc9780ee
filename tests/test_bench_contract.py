"""bench.py contract on the CPU: the reference arm (`--impl reference`) prints exactly one JSON line with the keys the
driver reads, on the CUDA arm's metric / unit / workload; without a GPU the CUDA arm must fail loudly (no CPU path)."""
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _run(*argv, env=None):
    return subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), *argv], capture_output=True, text=True,
                          timeout=300, cwd=ROOT, env={**os.environ, **(env or {})})


def test_reference_arm_line():
    p = _run("--impl", "reference", "--steps", "2", "--warmup", "1")
    assert p.returncode == 0, p.stderr[-2000:]
    lines = [x for x in p.stdout.splitlines() if x.strip()]
    assert len(lines) == 1                                            # ONE JSON line on stdout, nothing else
    d = json.loads(lines[0])
    assert d["impl"] == "reference" and d["metric"] == "env-steps/s incl. DeepMimic reward"
    assert d["unit"] == "env-steps/s" and d["higher_is_better"] is True and d["n_gpus"] == 1
    assert d["value"] > 100 and d["steps"] == 2 and d["gpu_launches"] == 0
    assert "BASELINE.json configs[1]" in d["config"]["workload"] and d["config"]["envs_per_gpu"] == 4096
    cb = d["cpu_baseline"]
    # "reference": the reference's own Python (baseline/_ref, made by baseline/build_ref.py) ran; "port": the numpy
    # restatement stood in because the copy is absent.  Either way the sample names the library that did the physics.
    have_ref = os.path.isdir(os.path.join(ROOT, "baseline", "_ref", "drloco")) or os.path.isdir("/root/reference/drloco")
    assert cb["kind"] == ("reference" if have_ref else "port")
    assert cb["cores"] >= 1 and cb["value"] == d["value"] and "liboracle.so" in cb["sample"]
    assert 0.0 < cb["physics_fraction"] < 1.0
    assert d["e2e"] == {"value": d["value"], "unit": "env-steps/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}


def test_reference_arm_other_ranks_exit_quietly():
    p = _run("--impl", "reference", "--steps", "1", "--warmup", "1", "--gpus", "2", env={"RANK": "1", "WORLD_SIZE": "2"})
    assert p.returncode == 0 and p.stdout.strip() == ""


def test_cuda_arm_fails_loudly_without_a_gpu():
    import torch
    if torch.cuda.is_available():
        return
    p = _run("--steps", "1", "--warmup", "1", "--no-cpu-baseline")
    assert p.returncode != 0 and p.stdout.strip() == ""               # no number is printed from a CPU fallback
