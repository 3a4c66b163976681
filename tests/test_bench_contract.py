"""CPU tier: the reference arm of bench.py (the one arm that runs without a GPU) prints the contract's JSON line, and the GPU
arm refuses to run - loudly - where there is no B200 (no CPU fallback behind the headline number)."""
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
BASELINE = json.load(open(os.path.join(ROOT, "BASELINE.json")))


def run_bench(*args, env=None):
    e = dict(os.environ)
    e.update(env or {})
    return subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), *args], capture_output=True, text=True, timeout=600, cwd=ROOT, env=e)


def test_reference_arm_prints_the_contract_line():
    r = run_bench("--impl", "reference", "--steps", "1", "--warmup", "1")
    assert r.returncode == 0, r.stderr[-2000:]
    lines = [l for l in r.stdout.splitlines() if l.startswith("{")]
    assert len(lines) == 1, r.stdout
    d = json.loads(lines[0])
    assert d["impl"] == "reference" and d["n_gpus"] == 1 and d["steps"] == 1 and d["warmup"] == 1
    assert d["unit"] == "Gpairs/s" and d["higher_is_better"] is True and d["vs_baseline"] is None
    assert d["metric"].split(";")[0].strip().startswith(BASELINE["metric"].split(";")[0].split(" match")[0])
    assert d["value"] > 0 and d["ms_per_step"] > 0 and d["gpu_launches"] == 0
    assert d["data"] == "synthetic" and d["config"]["workload"].startswith("C4")
    # its own numbers: nothing crosses PCIe, the cpu_baseline object describes this very run
    assert d["e2e"] == {"value": d["value"], "unit": d["unit"], "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}
    cb = d["cpu_baseline"]
    assert cb["value"] == d["value"] and cb["cores"] >= 1 and cb["kind"] in ("reference", "port") and cb["sample"]


def test_reference_arm_other_ranks_exit_quietly():
    # under torchrun only rank 0 runs the CPU matcher; the others leave at once with status 0 and print nothing
    r = run_bench("--impl", "reference", "--gpus", "2", "--steps", "1", "--warmup", "1", env={"RANK": "1", "LOCAL_RANK": "1", "WORLD_SIZE": "2"})
    assert r.returncode == 0, r.stderr[-2000:]
    assert not [l for l in r.stdout.splitlines() if l.startswith("{")]


def test_gpu_arm_fails_loudly_without_a_gpu():
    import torch
    if torch.cuda.is_available():
        import pytest
        pytest.skip("this box has a GPU")
    r = run_bench("--steps", "1", "--warmup", "1", "--no-extra", "--no-cpu-baseline")
    assert r.returncode != 0
    assert not [l for l in r.stdout.splitlines() if l.startswith("{")], "a bench line without a GPU would be a CPU fallback"
