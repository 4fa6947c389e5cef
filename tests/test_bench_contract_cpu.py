"""bench.py's reference arm (the CPU restatement of the reference's sparse mul! loops) runs without a GPU and prints
ONE JSON line with the keys the driver reads; the product arm must refuse to run without a device (no CPU fallback)."""
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def run(*args):
    return subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), *args], capture_output=True, text=True,
                          cwd=ROOT, timeout=300)


def test_reference_arm_line():
    r = run("--impl", "reference", "--steps", "1", "--warmup", "0", "--n-ref", "12")
    assert r.returncode == 0, r.stderr[-2000:]
    lines = [ln for ln in r.stdout.splitlines() if ln.strip()]
    assert len(lines) == 1
    d = json.loads(lines[0])
    assert d["impl"] == "reference" and d["higher_is_better"] is True and d["n_gpus"] == 1
    assert d["unit"] == "DOF/s" and d["value"] > 0 and d["warmup"] >= 3 and d["dtype"] == "f64"
    assert d["cpu_baseline"]["kind"] in ("port", "reference") and d["cpu_baseline"]["cores"] >= 1
    assert d["cpu_baseline"]["value"] == d["value"] and "sample" in d["cpu_baseline"]
    assert d["e2e"] == {"value": d["value"], "unit": d["unit"], "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}
    assert "workload" in d["config"] and "model" not in d["config"]


def test_product_arm_needs_a_gpu():
    import torch
    if torch.cuda.is_available():
        import pytest
        pytest.skip("a GPU is present: the product arm runs (covered by the gpu tests and the bench itself)")
    r = run("--steps", "1", "--warmup", "0", "--cells", "4", "--no-e2e", "--no-cpu", "--no-extras", "--no-wave2d")
    assert r.returncode != 0                       # fails loudly: libddf_b200 has no CPU path
    assert not [ln for ln in r.stdout.splitlines() if ln.startswith("{")]


def test_reference_arm_under_torchrun_prints_once():
    """The driver launches both arms the same way (torchrun for N > 1): rank 0 alone runs and prints the reference
    line, the other ranks exit 0 without work."""
    r = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2",
                        "--master-addr", "127.0.0.1", "--master-port", "29541", os.path.join(ROOT, "bench.py"),
                        "--impl", "reference", "--gpus", "2", "--steps", "1", "--warmup", "0", "--n-ref", "12"],
                       capture_output=True, text=True, cwd=ROOT, timeout=300)
    assert r.returncode == 0, r.stderr[-2000:]
    lines = [ln for ln in r.stdout.splitlines() if ln.startswith("{")]
    assert len(lines) == 1 and json.loads(lines[0])["impl"] == "reference"
