#!/usr/bin/env python
"""Run under torchrun: every rank reduces its share of the rows of each step, D' rows are all-gathered
over NCCL, the echelon is replicated; every rank's result must equal the oracle's rows bit for bit.
usage: torchrun --nproc-per-node N tools/check_multi.py DIR   (DIR from `gamba_oracle --dump-steps`)"""
import glob
import os
import sys

import torch
import torch.distributed as dist

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from gamba_b200 import LinalgEngine, load_rows, load_step, parallel  # noqa: E402


def main():
    d = sys.argv[1]
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    gather = parallel.make_allgather()
    bcast = parallel.make_broadcast() if os.environ.get("GAMBA_BCAST", "1") == "1" else None
    bad = n = 0
    engines = {}
    for f in sorted(glob.glob(os.path.join(d, "step_*.bin"))):
        s = load_step(f)
        want = load_rows(f.replace("step_", "rows_"))
        e = engines.get(s.p)
        if e is None:
            e = engines[s.p] = LinalgEngine(s.p, device=local, rank=rank, world_size=world)
            e.set_allgather(gather)
            if bcast:
                e.set_broadcast(bcast)
            for kv in filter(None, os.environ.get("CHECK_MULTI_OPTS", "").split(",")):   # e.g. schur_min_top=1,block=2
                k, v = kv.split("=")
                e.set_option(k, int(v))
        # with the broadcast only rank 0 hands the step over; the others receive it over NCCL
        got = e.reduce(s if (rank == 0 or not bcast) else None, final=f.endswith("step_final.bin"))
        n += 1
        bad += not got.same_as(want)
    t = torch.tensor([bad, n], device="cuda")
    dist.all_reduce(t)
    if rank == 0:
        print(f"multi-gpu check: world={world} steps={n} rank-steps differing={int(t[0])} of {int(t[1])}")
    dist.destroy_process_group()
    sys.exit(1 if int(t[0]) else 0)


if __name__ == "__main__":
    main()
