"""N>1 on real GPUs (skipped on a 1-GPU box): rows of C|D sharded r % 2 over two ranks, D' all-gathered over
NCCL, echelon replicated — every rank's result bit-exact against the oracle."""
import os
import subprocess
import sys

import pytest

from conftest import ROOT

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("name,opts", [("cyclic7-31", ""), ("cyclic7-31", "schur_min_top=1,block=2"),
                                       ("cyclic8-15", "schur_min_top=1,block=2"), ("cyclic8-15", "schur_min_top=1,block=0"),
                                       ("cyclic8-15", "schur_min_top=1,panel=2,panel_t=128,mc=2"),
                                       ("cyclic8-31", "schur_min_top=1,panel=0,block=2,mc=2")])
def test_two_ranks_match_oracle(step_dirs, name, opts):
    """Default path, and the tensor-core product with the blocked / row-per-CTA sparse solve on sharded rows; mc=2: Monte-Carlo
    combinations of the rows (each rank forms and eliminates its share of them; draws that miss the rank margin are topped up
    behind the gathered block of the earlier ones) in front of either blocked solve."""
    import torch
    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs")
    d = step_dirs(name) if name == "cyclic7-31" else step_dirs(name, ("--dump-min-nnz", "100000"))
    r = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2",
                        "--master-addr", "127.0.0.1", "--master-port", "29533", os.path.join(ROOT, "tools", "check_multi.py"), d],
                       capture_output=True, text=True, timeout=600, env=dict(os.environ, CHECK_MULTI_OPTS=opts))
    assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-2000:]
    assert "differing=0" in r.stdout


@pytest.mark.parametrize("name", ["cyclic7-31", "kat10-15", "eco10-31"])
def test_product_driver_on_two_gpus_matches_golden(built, tmp_path, name):
    """`gamba --gpus 2`: one process, one engine per GPU, rows of C|D sharded r % 2, the step replicated with ncclBroadcast and
    D' gathered with ncclAllGather (the C++ wiring of the two ABI callbacks, gamba_b200/host/engine_cuda.cpp)."""
    import torch
    from conftest import GOLD_IN, golden_bytes
    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs")
    out = tmp_path / "out.txt"
    r = subprocess.run([os.path.join(ROOT, "gamba_b200", "bin", "gamba"), "-i", os.path.join(GOLD_IN, name + ".txt"), "-o", str(out),
                        "-v", "1", "--gpus", "2"], capture_output=True, text=True, timeout=900)
    assert r.returncode == 0, r.stderr[-2000:]
    assert "engine = cuda-b200-multi" in r.stdout
    assert out.read_bytes() == golden_bytes(name)
