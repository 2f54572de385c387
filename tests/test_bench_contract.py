"""bench.py's reference arm and output contract, on CPU: `--impl reference` times the NumPy port of the reference graph (the
oracle, the only CPU implementation there is: TensorFlow 1.x cannot run here) and prints exactly ONE JSON line on stdout;
under torchrun only rank 0 works; the B200 arm refuses to run without a GPU instead of falling back."""
import json
import os
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def run(args, env=None):
    e = dict(os.environ, **(env or {}))
    return subprocess.run([sys.executable, os.path.join(ROOT, "bench.py")] + args, capture_output=True, text=True, cwd=ROOT, env=e,
                          timeout=600)


def test_reference_arm_prints_one_json_line_with_the_contract_keys():
    r = run(["--impl", "reference", "--steps", "1", "--warmup", "1"])
    assert r.returncode == 0, r.stderr[-2000:]
    lines = [l for l in r.stdout.splitlines() if l.strip()]
    assert len(lines) == 1
    d = json.loads(lines[0])
    assert d["impl"] == "reference" and d["metric"] == "bilstm_train_captions_per_sec" and d["unit"] == "captions/s"
    assert d["higher_is_better"] is True and d["n_gpus"] == 1 and d["steps"] == 1 and d["vs_baseline"] is None
    assert d["config"]["workload"] == "card2048" and d["value"] > 0
    assert d["cpu_baseline"]["kind"] == "port" and d["cpu_baseline"]["cores"] >= 1 and d["cpu_baseline"]["value"] == d["value"]
    assert d["e2e"] == dict(value=d["value"], unit="captions/s", h2d_bytes_per_step=0, d2h_bytes_per_step=0)
    assert d["gpu_launches"] == 0


def test_reference_arm_uses_every_host_core_and_the_b200_arms_batch_under_torchrun():
    """torchrun exports OMP_NUM_THREADS=1: the CPU arm must lift that limit itself (round 1's N>1 ratios were taken against a
    single-threaded BLAS), and it runs the same batch per step as the B200 arm (2048 captions of card2048)."""
    r = run(["--impl", "reference", "--gpus", "2", "--steps", "1", "--warmup", "1"],
            env=dict(RANK="0", WORLD_SIZE="2", LOCAL_RANK="0", OMP_NUM_THREADS="1"))
    assert r.returncode == 0, r.stderr[-2000:]
    d = json.loads(r.stdout.strip())
    assert d["cpu_baseline"]["cores"] == len(os.sched_getaffinity(0))
    assert d["config"]["batch_per_step"] == 2048 and d["config"]["batch_per_gpu"] == 2048


def test_reference_arm_runs_on_rank_zero_only():
    r = run(["--impl", "reference", "--gpus", "2", "--steps", "1", "--warmup", "1"], env=dict(RANK="1", WORLD_SIZE="2", LOCAL_RANK="1"))
    assert r.returncode == 0 and r.stdout.strip() == ""


def test_b200_arm_has_no_cpu_fallback():
    import torch
    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    r = run(["--steps", "1", "--warmup", "1"])
    assert r.returncode != 0 and r.stdout.strip() == "" and "no CPU fallback" in r.stderr
