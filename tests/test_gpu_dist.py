"""Multi-GPU test (needs >= 2 GPUs on the box, otherwise skipped): the domain-decomposed integrator must
reproduce the single-GPU trajectory bit for bit (same colours, same arithmetic, same summation order)."""
import os
import subprocess
import sys

import pytest
import torch

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.mark.skipif(torch.cuda.device_count() < 2, reason="needs two GPUs")
@pytest.mark.parametrize("dataflow", ["1", "0"])
def test_domain_decomposition_two_gpus_bitwise(dataflow):
    """Barrier-free (default) and with colour barriers + epochs (VBDX_DATAFLOW=0): both must equal the single-GPU run bit
    for bit, hence each other."""
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2", "--master-addr", "127.0.0.1",
           "--master-port", "29541" if dataflow == "1" else "29543", os.path.join(ROOT, "tools", "dist_check.py"), "10", "5"]
    out = subprocess.run(cmd, capture_output=True, text=True, timeout=300, cwd=ROOT, env=dict(os.environ, VBDX_DATAFLOW=dataflow))
    assert out.returncode == 0, out.stderr[-2000:]
    lines = [l for l in out.stdout.splitlines() if l.startswith("world=2")]
    assert len(lines) == 4, out.stdout
    for l in lines:
        assert "rel L2 vs single GPU = 0.000e+00" in l, l
