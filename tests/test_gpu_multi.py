"""N>1 on real GPUs (skipped on a 1-GPU box): rows of C|D sharded r % 2 over two ranks, D' all-gathered over
NCCL, echelon replicated — every rank's result bit-exact against the oracle."""
import os
import subprocess
import sys

import pytest

from conftest import ROOT

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("name,opts", [("cyclic7-31", ""), ("cyclic7-31", "schur_min_top=1,block=2"),
                                       ("cyclic8-15", "schur_min_top=1,block=2"), ("cyclic8-15", "schur_min_top=1,block=0")])
def test_two_ranks_match_oracle(step_dirs, name, opts):
    """Default path, and the tensor-core product with the blocked / row-per-CTA sparse solve on sharded rows."""
    import torch
    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs")
    d = step_dirs(name) if name == "cyclic7-31" else step_dirs(name, ("--dump-min-nnz", "100000"))
    r = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2",
                        "--master-addr", "127.0.0.1", "--master-port", "29533", os.path.join(ROOT, "tools", "check_multi.py"), d],
                       capture_output=True, text=True, timeout=600, env=dict(os.environ, CHECK_MULTI_OPTS=opts))
    assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-2000:]
    assert "differing=0" in r.stdout
