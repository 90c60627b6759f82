"""bench.py's reference arm runs without a GPU: its one JSON line carries the keys the driver reads.
(The CUDA arm is exercised on the GPU box; here its argument handling and the refusal to run
without a device are checked.)"""
import json
import os
import subprocess
import sys

from tests.helpers import ROOT


def _run(args, timeout=600):
    env = dict(os.environ, PYTHONPATH=ROOT)
    return subprocess.run([sys.executable, os.path.join(ROOT, "bench.py")] + args, capture_output=True, text=True,
                          timeout=timeout, env=env, cwd=ROOT)


def test_reference_arm_prints_one_json_line():
    r = _run(["--impl", "reference", "--steps", "1", "--warmup", "0", "--workload", "brachistochrone20"])
    assert r.returncode == 0, r.stderr[-2000:]
    lines = [l for l in r.stdout.splitlines() if l.strip()]
    assert len(lines) == 1
    d = json.loads(lines[0])
    assert d["impl"] == "reference" and d["metric"] == "collocation defect+FD-Jacobian evals/sec"
    assert d["unit"] == "evals/s" and d["higher_is_better"] is True and d["value"] > 0
    assert d["n_gpus"] == 1 and d["steps"] == 1 and d["scaling"] == "weak" and d["dtype"] == "f64"
    cb = d["cpu_baseline"]
    from oracle import ref_loader
    # the real reference wherever it can be imported (/root/reference here, baseline/_ref on the GPU box)
    assert cb["kind"] == ("reference" if ref_loader.reference_available() else "port")
    assert cb["cores"] >= 1 and cb["value"] == d["value"] and "sample" in cb
    assert d["e2e"] == {"value": d["value"], "unit": "evals/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}
    assert d["config"]["workload"].startswith("brachistochrone20")


def test_reference_arm_falls_back_to_the_oracle_port():
    env = dict(os.environ, PYTHONPATH=ROOT, OGB200_CPU_KIND="port")
    r = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--steps", "1", "--warmup", "0",
                        "--workload", "brachistochrone20"], capture_output=True, text=True, timeout=600, env=env, cwd=ROOT)
    assert r.returncode == 0, r.stderr[-2000:]
    d = json.loads(r.stdout.strip().splitlines()[-1])
    assert d["cpu_baseline"]["kind"] == "port" and d["value"] > 0


def test_reference_arm_other_ranks_do_no_work():
    env = dict(os.environ, PYTHONPATH=ROOT, RANK="1", WORLD_SIZE="2", LOCAL_RANK="1")
    r = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--gpus", "2",
                        "--steps", "1", "--warmup", "0"], capture_output=True, text=True, timeout=300, env=env, cwd=ROOT)
    assert r.returncode == 0 and r.stdout.strip() == ""


def test_cuda_arm_refuses_to_run_without_a_gpu():
    import torch
    if torch.cuda.is_available():
        import pytest
        pytest.skip("a GPU is present")
    r = _run(["--steps", "1", "--no-cpu-baseline", "--workload", "brachistochrone20"])
    assert r.returncode != 0 and "CUDA" in (r.stderr + r.stdout)
