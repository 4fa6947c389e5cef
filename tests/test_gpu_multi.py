"""Multi-GPU parity inside the tested tier: self-spawns one rank per visible GPU (at most 4)
with torch.distributed.run and runs tools/dist_check.py -- the slab-decomposed fused wave step (peer-memory halo AND
the NCCL fallback), the fused Poisson sweep, Tsit5 with ghost exchange, the rank-R ghost exchange and the sharded
d_k / star_k apply with global numbering, every one compared BIT FOR BIT with the undivided mesh on one GPU, from
random (non-symmetric, non-zero on the interface planes) fields.  Skipped on a box with fewer than 2 GPUs."""
import os
import subprocess
import sys

import pytest

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_slab_decomposition_bit_exact_on_all_visible_gpus():
    import torch
    ngpu = torch.cuda.device_count()
    if ngpu < 2:
        pytest.skip(f"{ngpu} GPU visible: the multi-GPU parity check needs at least 2 (run under gpurun --gpus 2)")
    world = min(ngpu, 4)
    env = dict(os.environ, MASTER_ADDR="127.0.0.1")
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={world}",
           "--master-addr", "127.0.0.1", "--master-port", "29533", os.path.join(ROOT, "tools", "dist_check.py")]
    res = subprocess.run(cmd, cwd=ROOT, env=env, capture_output=True, text=True, timeout=900)
    out = res.stdout + res.stderr
    lines = [ln for ln in out.splitlines() if ln.startswith("dist_check")]
    assert res.returncode == 0, out[-4000:]
    assert len(lines) >= 30 and all("BIT-EXACT" in ln for ln in lines), "\n".join(lines)
    assert sum("general partition" in ln for ln in lines) == 2, "\n".join(lines)
    assert any("halo=p2p" in ln for ln in lines) and any("halo=nccl" in ln for ln in lines), "\n".join(lines)
    t5 = [ln for ln in lines if ln.startswith("dist_check tsit5")]
    assert any("halo=p2p" in ln for ln in t5) and any("halo=nccl" in ln for ln in t5), "\n".join(t5)
