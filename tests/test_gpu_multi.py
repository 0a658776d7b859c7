"""N>1 on real GPUs (skipped on a 1-GPU box): rows of C|D sharded r % 2 over two ranks, D' all-gathered over
NCCL, echelon replicated — every rank's result bit-exact against the oracle."""
import os
import subprocess
import sys

import pytest

from conftest import ROOT

pytestmark = pytest.mark.gpu


def test_two_ranks_match_oracle(step_dirs):
    import torch
    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs")
    d = step_dirs("cyclic7-31")
    r = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2",
                        "--master-addr", "127.0.0.1", "--master-port", "29533", os.path.join(ROOT, "tools", "check_multi.py"), d],
                       capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-2000:]
    assert "differing=0" in r.stdout
