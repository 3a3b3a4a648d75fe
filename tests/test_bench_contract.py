"""bench.py's output contract, checked where no GPU is needed: the reference arm (`--impl reference`) runs the CPU
implementation of the path and must put exactly ONE JSON line on stdout -- everything else (library chatter such as
NCCL's version line in multi-GPU runs, warnings) goes to stderr -- with the keys the driver reads."""
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def run_bench(args, env=None):
    e = dict(os.environ)
    e.update(env or {})
    return subprocess.run([sys.executable, os.path.join(ROOT, "bench.py")] + args, cwd=ROOT, env=e,
                          capture_output=True, text=True, timeout=600)


def test_reference_arm_prints_one_json_line():
    p = run_bench(["--impl", "reference", "--steps", "1", "--warmup", "0"])
    assert p.returncode == 0, p.stderr[-2000:]
    lines = [ln for ln in p.stdout.splitlines() if ln.strip()]
    assert len(lines) == 1, p.stdout[:2000]
    d = json.loads(lines[0])
    assert d["impl"] == "reference" and d["unit"] == "solves/s" and d["higher_is_better"] is True
    assert d["value"] > 0 and d["n_gpus"] == 1 and d["steps"] == 1
    assert d["cpu_baseline"]["kind"] in ("reference", "port") and d["cpu_baseline"]["cores"] >= 1
    assert d["cpu_baseline"]["value"] == d["value"]
    assert d["e2e"] == {"value": d["value"], "unit": d["unit"], "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}
    assert "workload" in d["config"] and "sample" in d["config"]["workload"]
    # the reference arm runs without the product: no CUDA library of this repository in the process
    assert all("mpc_b200" not in x for x in d["loaded_repo_libraries"]), d["loaded_repo_libraries"]
    assert any(x.startswith("oracle") for x in d["loaded_repo_libraries"])


def test_reference_arm_other_ranks_do_no_work():
    """Under torchrun (N > 1) rank 0 alone runs the reference arm; the other ranks exit 0 and print nothing."""
    p = run_bench(["--impl", "reference", "--gpus", "2", "--steps", "1", "--warmup", "0"],
                  env={"RANK": "1", "LOCAL_RANK": "1", "WORLD_SIZE": "2"})
    assert p.returncode == 0, p.stderr[-2000:]
    assert p.stdout.strip() == ""


def test_our_arm_fails_loudly_without_a_gpu():
    """No CPU fallback: without a CUDA device the product arm must refuse, not fall back to the oracle."""
    import torch
    if torch.cuda.is_available():
        import pytest
        pytest.skip("a GPU is present")
    p = run_bench(["--steps", "1", "--warmup", "3"])
    assert p.returncode != 0
    assert p.stdout.strip() == ""
    assert "no CUDA device" in p.stderr
