#!/usr/bin/env python
"""Benchmark of the hot path: F4 linear-algebra throughput, nnz reduced per second.

Workload (config.workload): the F4 step matrices of cyclic9-15 (BASELINE.json: the metric is quoted
on cyclic9 / kat15; cyclic9-15 is the largest named system whose whole run fits the default time
budget) with at least --min-nnz non-zeros. The matrices are produced on the box by the product
driver (`gamba --dump-steps`, untimed workload generation). One bench "step" = one pass of
gamba_cuda_reduce over that batch of matrices.

  value   Σ nnz_in / time, steps already staged in HBM (gamba_cuda_stage once, untimed;
          gamba_cuda_reduce_staged in the timed region: kernels + result D2H)
  e2e     the same through gamba_cuda_reduce with pinned HOST buffers: H2D + staging kernels +
          elimination + echelon + D2H all inside the timed region
  roofline  elimination stage (C|D by A|B = elim_kernel for the multipliers + the tensor-core product
          D' = D + M B): algorithmic bytes (fma_top*(4+w) + nnz_in*(4+w) + n_bot*n_nonpiv*8)
          / CUDA-event time of the stage / measured HBM peak; `parts` splits the time and gives the
          product's int8 multiply-add rate against the measured IMMA peak
  cpu_baseline  the CPU oracle (a port: the reference's engine is closed source) on a bounded sample

`--impl reference` times the oracle port on bounded samples of the same workload (1 thread: the
reference is single-threaded, README.md:9,71).
"""
from __future__ import annotations

import argparse
import glob
import json
import os
import shutil
import subprocess
import sys
import tempfile
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

import numpy as np  # noqa: E402

METRIC = "F4 lin-alg nnz reduced/s"
WORK_DIR = os.path.join(tempfile.gettempdir(), "gamba_b200_bench")


def measured_peaks():
    try:
        with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as f:
            return float(json.load(f)["hbm_gbs"]), "measured (MEASURED_PEAKS.json)"
    except Exception:  # noqa: BLE001
        return 6650.0, "fallback (B200_PROFILING.md)"


def measured_traffic():
    """DRAM bytes per elimination-kernel launch from the committed ncu capture (profiles/), or None."""
    try:
        with open(os.path.join(ROOT, "profiles", "r01d_elim_traffic.json")) as f:
            return float(json.load(f)["mean_dram_bytes_per_launch"])
    except Exception:  # noqa: BLE001
        return None


def cpu_model() -> str:
    try:
        for line in open("/proc/cpuinfo"):
            if line.startswith("model name"):
                return line.split(":", 1)[1].strip()
    except OSError:
        pass
    return "unknown"


class ClockSampler:
    """nvidia-smi clocks / throttle reasons during the timed region (B200_PROFILING.md)."""

    def __init__(self, index: int):
        self.index, self.rows, self.proc = index, [], None

    def __enter__(self):
        q = ("clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
             "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.index), f"--query-gpu={q}",
                                          "--format=csv,noheader,nounits", "-lms", "200"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.t = threading.Thread(target=self._read, daemon=True)
            self.t.start()
        except OSError:
            self.proc = None
        return self

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([x.strip() for x in line.split(",")])

    def __exit__(self, *exc):
        if self.proc:
            self.proc.terminate()
            try:
                self.proc.wait(timeout=5)
            except subprocess.TimeoutExpired:
                self.proc.kill()

    def summary(self):
        sm, mx, reasons = [], 0, set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for r in self.rows:
            try:
                sm.append(float(r[0])); mx = max(mx, float(r[1]))
            except (ValueError, IndexError):
                continue
            for n, v in zip(names, r[2:6]):
                if v.lower().startswith("active"):
                    reasons.add(n)
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": mx or None, "reasons": sorted(reasons),
                "samples": len(sm)}


def generate_workload(system: str, min_nnz: int, rank: int, world: int, max_spairs: int = 2000) -> list[str]:
    """Step matrices of `system` with >= min_nnz non-zeros, produced by the product driver on this box."""
    from gamba_b200 import systems
    d = os.path.join(WORK_DIR, f"{system}_{min_nnz}" + ("" if max_spairs == 2000 else f"_sp{max_spairs}"))
    done = os.path.join(d, "DONE")
    if rank == 0 and not os.path.exists(done):
        shutil.rmtree(d, ignore_errors=True)
        os.makedirs(d)
        inp = systems.write_system(system, os.path.join(d, system + ".txt"))
        t = time.time()
        r = subprocess.run([os.path.join(ROOT, "gamba_b200", "bin", "gamba"), "-i", inp, "-o", os.path.join(d, "out.txt"),
                            "-v", "1", "--dump-steps", d, "--dump-min-nnz", str(min_nnz), "--max-spairs", str(max_spairs)],
                           capture_output=True, text=True)
        if r.returncode != 0:
            raise RuntimeError("workload generation failed (no CPU fallback): " + r.stderr[-2000:])
        with open(os.path.join(d, "driver.log"), "w") as f:
            f.write(r.stdout)
        with open(done, "w") as f:
            f.write(f"{time.time() - t:.1f}\n")
    if world > 1:
        import torch.distributed as dist
        dist.barrier()
    files = sorted(f for f in glob.glob(os.path.join(d, "step_*.bin")) if not f.endswith("step_final.bin"))
    if not files:
        raise RuntimeError("no step matrices generated")
    return files


def load_workload(args, rank: int, world: int):
    """List of step matrices: the named system's F4 steps, or `synthetic:ROWSxCOLS[:density[:bits]]`
    (BASELINE.json configs[4]: F4-shaped random matrix, every rank generates the same one from the seed)."""
    from gamba_b200 import load_step
    from gamba_b200.stepfile import synthetic_step_large
    if args.system.startswith("synthetic:"):
        parts = args.system.split(":")
        nr, nc = (int(x) for x in parts[1].lower().split("x"))
        dens = float(parts[2]) if len(parts) > 2 else 0.005
        p = {"15": 32003, "31": 2147483647}[parts[3] if len(parts) > 3 else "31"]
        return [synthetic_step_large(nr, nc, dens, p)], f"synthetic F4-shaped matrix {nr} x {nc}, density {dens}, p={p}, seed 967557673"
    files = generate_workload(args.system, args.min_nnz, rank, world, args.max_spairs)
    return [load_step(f) for f in files], (f"{args.system} F4 step matrices with >= {args.min_nnz} nnz" +
                                           ("" if args.max_spairs == 2000 else f" (--max-spairs {args.max_spairs})"))


def bytes_alg(step, fma_top: int) -> float:
    w = 2 if step.p < 65536 else 4
    return fma_top * (4 + w) + step.nnz * (4 + w) + step.n_bot * step.n_nonpiv * 8


def subsample_rows(step, every: int):
    """Bounded sample of one step: all pivot rows, every `every`-th row to reduce."""
    from gamba_b200.stepfile import Step
    if every <= 1:
        return step, 1.0
    nt = step.n_top
    keep = np.arange(0, step.n_bot, every) + nt
    rp = step.row_ptr.astype(np.int64)
    parts = [step.col[: rp[nt]]]
    lens = []
    for r in keep:
        parts.append(step.col[rp[r]: rp[r + 1]]); lens.append(rp[r + 1] - rp[r])
    row_ptr = np.concatenate([rp[: nt + 1], rp[nt] + np.cumsum(lens)]).astype(np.uint64)
    s = Step(step.p, nt, len(keep), step.n_cols, row_ptr, np.concatenate(parts),
             np.concatenate([step.coeff_id[:nt], step.coeff_id[keep]]), step.coeff_len, step.pool)
    return s, len(keep) / step.n_bot


def run_reference(args, rank: int, world: int):
    """Reference arm: the CPU oracle (port of the closed engine) on bounded samples of the workload."""
    if rank != 0:
        return
    from gamba_b200 import load_step
    from oracle import oracle
    oracle.build()
    steps, wl_desc = load_workload(args, 0, 1)
    # smallest matrix of the batch, thinned so that one bench step is a few seconds of CPU work
    s = min(steps, key=lambda x: x.nnz * x.n_bot)
    est_fma = 0.5 * s.n_bot * (s.nnz / s.n_rows) * s.n_top * 0.25   # rough: quarter of the pivots hit per row
    every = max(1, int(est_fma / 8e8 / args.ref_seconds))
    sample, frac = subsample_rows(s, every)
    nnz_sample = int(s.row_ptr[s.n_top]) * frac + (sample.nnz - int(s.row_ptr[s.n_top]))
    for _ in range(args.warmup if args.warmup < 2 else 1):
        oracle.reduce(sample)
    t0 = time.perf_counter()
    secs = 0.0
    for _ in range(args.steps):
        _, _, sec = oracle.reduce(sample)
        secs += sec
    wall = time.perf_counter() - t0
    value = nnz_sample * args.steps / secs
    desc = (f"{args.system} F4 step {s.n_rows}x{s.n_cols} ({s.nnz} nnz), all pivot rows + every {every}-th of the "
            f"{s.n_bot} rows to reduce; nnz counted = that fraction of nnz_in")
    print(json.dumps({
        "impl": "reference", "metric": METRIC, "value": value, "unit": "nnz/s", "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": 1e3 * wall / args.steps, "higher_is_better": True, "scaling": "strong",
        "vs_baseline": None, "dtype": "u64 (CPU oracle accumulators) over F_p", "data": "synthetic (generated system, no dataset)",
        "config": {"workload": wl_desc, "sample": desc},
        "cpu_baseline": {"value": value, "unit": "nnz/s", "cores": 1, "kind": "port", "sample": desc, "cpu": cpu_model(),
                         "note": "oracle port; the reference's own engine (linalg_v3.hpp, kernel/*) is closed source"},
        "e2e": {"value": value, "unit": "nnz/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }))


def run_ours(args, rank: int, world: int, local_rank: int):
    import torch
    from gamba_b200 import LinalgEngine, load_step, parallel
    from gamba_b200.stepfile import Step
    if not torch.cuda.is_available():
        raise RuntimeError("bench.py needs a CUDA device: the product has no CPU path")
    torch.cuda.set_device(local_rank)
    if world > 1:
        import torch.distributed as dist
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    steps, wl_desc = load_workload(args, rank, world)
    if args.max_matrices:
        steps = sorted(steps, key=lambda s: -s.nnz)[: args.max_matrices]
    total_nnz = sum(s.nnz for s in steps)

    def pinned(s: Step) -> Step:  # host buffers the e2e leg copies from
        def pin(a):
            t = torch.from_numpy(np.ascontiguousarray(a)).pin_memory()
            return t.numpy(), t
        keep = []
        out = []
        for a in (s.row_ptr, s.col, s.coeff_id, s.coeff_len, s.pool):
            n, t = pin(a); out.append(n); keep.append(t)
        st = Step(s.p, s.n_top, s.n_bot, s.n_cols, *out)
        st._pins = keep
        return st
    steps = [pinned(s) for s in steps]

    engines = []
    gather = parallel.make_allgather() if world > 1 else None
    bcast = parallel.make_broadcast() if world > 1 and not args.no_bcast else None
    for s in steps:
        e = LinalgEngine(s.p, device=local_rank, rank=rank, world_size=world)
        if gather:
            e.set_allgather(gather)
        if bcast:
            e.set_broadcast(bcast)
        for kv in args.opt:
            k, v = kv.split("=")
            e.set_option(k, int(v))
        engines.append(e)

    def sync():
        torch.cuda.synchronize()
        if world > 1:
            import torch.distributed as dist
            dist.barrier()

    def launches():
        return sum(e.counters()["kernel_launches"] for e in engines)

    # ---- value: steps resident in HBM --------------------------------------------------------
    def mine(s):   # with the NCCL broadcast only rank 0 hands the step over
        return s if (rank == 0 or not bcast) else None
    for e, s in zip(engines, steps):
        e.stage(mine(s))
        for _ in range(3):     # untimed: the engine picks its kernels for a step from what the previous steps measured
            e.reduce_staged()  # (F4 steps change slowly: one step per mode, then the calibrated choice); here every
            e.stage(mine(s))   # matrix has its own engine, so it sees itself first
    for _ in range(args.warmup):
        for e in engines:
            e.reduce_staged()
    sync()
    elim_s = [0.0] * len(steps); ech_s = [0.0] * len(steps); dev_s = 0.0
    sparse_s = 0.0; schur_s = 0.0; schur_imma = 0.0
    fma_top = [0] * len(steps)
    l0 = launches()
    with ClockSampler(local_rank) as clk:
        sync()
        t0 = time.perf_counter()
        for _ in range(args.steps):
            for i, e in enumerate(engines):
                e.reduce_staged(copy=False)     # the result stays in the engine-owned host buffers the C ABI hands out
                elim_s[i] += e.last["seconds_eliminate"]; ech_s[i] += e.last["seconds_echelon"]
                dev_s += e.last["seconds_eliminate"] + e.last["seconds_echelon"]
                c_ = e.counters()
                fma_top[i] = c_["fma_top"]
                sparse_s += c_["seconds_sparse_last"]; schur_s += c_["seconds_schur_last"]
                schur_imma += c_["schur_macs"] * (4 if steps[i].p < 65536 else 16)   # limb products per mod-p multiply-add
        sync()
        t_res = time.perf_counter() - t0
    n_launch = launches() - l0
    clocks = clk.summary()

    # ---- e2e: pinned host buffers in, host buffers out -----------------------------------------
    for _ in range(max(1, args.warmup // 2)):
        for e, s in zip(engines, steps):
            e.reduce(mine(s))
    sync()
    t0 = time.perf_counter()
    h2d = d2h = 0
    for _ in range(args.steps):
        for e, s in zip(engines, steps):
            e.reduce(mine(s), copy=False)
            h2d += e.last["h2d_bytes"]; d2h += e.last["d2h_bytes"]
    sync()
    t_e2e = time.perf_counter() - t0

    if world > 1:
        import torch.distributed as dist
        tt = torch.tensor([t_res, t_e2e], device="cuda", dtype=torch.float64)
        dist.all_reduce(tt, op=dist.ReduceOp.MAX)
        t_res, t_e2e = float(tt[0]), float(tt[1])
        ft = torch.tensor(fma_top, device="cuda", dtype=torch.int64)
        dist.all_reduce(ft, op=dist.ReduceOp.SUM)   # every rank counted its own rows
        fma_top = [int(x) for x in ft]
        et = torch.tensor(elim_s, device="cuda", dtype=torch.float64)
        dist.all_reduce(et, op=dist.ReduceOp.MAX)
        elim_s = [float(x) for x in et]

    if rank == 0:
        peak, peak_src = measured_peaks()
        # dominant kernel: the C|D-by-A|B elimination; per launch = per matrix, averaged over the batch
        alg = sum(bytes_alg(s, f) for s, f in zip(steps, fma_top)) / len(steps)
        dur = sum(elim_s) / (args.steps * len(steps))
        achieved = alg / dur / 1e9 / world      # per GPU: every rank streams its share of the rows
        c = engines[0].counters()
        out = {
            "metric": METRIC, "value": total_nnz * args.steps / t_res, "unit": "nnz/s", "n_gpus": world,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * t_res / args.steps,
            "higher_is_better": True, "scaling": "strong", "vs_baseline": None,
            "dtype": "u32 residues mod p=%d (32-bit delayed-reduction lanes in the sparse solve; u8 limbs -> s32 on the tensor cores in the Schur product and the echelon)" % steps[0].p, "data": "synthetic (generated system, no dataset)",
            "config": {"workload": f"{wl_desc} ({len(steps)} matrices, "
                                   f"{total_nnz} nnz per pass, largest {max(s.n_rows for s in steps)} x {max(s.n_cols for s in steps)})",
                       "parallelism": (f"C|D rows sharded r % {world}, step uploaded by rank 0 and replicated by NCCL broadcast, "
                                       f"NCCL all-gather of D' rows, echelon replicated") if world > 1 else "1 GPU",
                       "l2": "inputs larger than L2 (every pass cycles through all matrices; smallest staged step > 126 MB)",
                       "max_spairs": args.max_spairs, "tile_cols": c["tile_cols"], "acc_bits": c["acc_bits"]},
            "e2e": {"value": total_nnz * args.steps / t_e2e, "unit": "nnz/s", "h2d_bytes_per_step": h2d // args.steps,
                    "d2h_bytes_per_step": d2h // args.steps, "ms_per_step": 1e3 * t_e2e / args.steps},
            "gpu_launches": n_launch,
            "clocks": clocks,
            "roofline": {"kernel": "elimination stage C|D by A|B: elim_kernel (walk of the pivot windows) + schur_tc_kernel per column "
                                   "tile (earlier pivots' contribution, tcgen05) for the multipliers M; schur_bt_kernel + schur_tc_kernel "
                                   "for D' = D + M B",
                         "bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s",
                         "frac": achieved / peak, "traffic": measured_traffic() if args.system == "cyclic9-15" else None,
                         "traffic_source": "ncu dram__bytes_read+write of the stage's kernels per matrix, profiles/r01d_elim_traffic.json",
                         "peak_source": peak_src,
                         "parts": {"sparse_solve_ms_per_step": 1e3 * sparse_s / args.steps, "schur_ms_per_step": 1e3 * schur_s / args.steps,
                                   "schur_tensor": {"achieved": schur_imma / max(schur_s, 1e-12) / 1e12, "peak": 2250.0, "unit": "int8 TMAC/s",
                                                    "frac": schur_imma / max(schur_s, 1e-12) / 1e12 / 2250.0,
                                                    "peak_source": "nominal dense int8 of tcgen05 on B200 (4.5 POP/s = 2x the bf16 rate; MEASURED_PEAKS.json "
                                                                   "has cuBLAS bf16 at 1648 TFLOP/s = 1648 int8-equivalent TMAC/s); the time includes "
                                                                   "zeroing the multiplier planes"}},
                         "bytes_alg_per_launch": alg, "avg_launch_ms": dur * 1e3,
                         "share_of_device_time": sum(elim_s) / max(dev_s, 1e-12),
                         "fma_top_per_pass": int(sum(fma_top))},
            "breakdown_ms_per_step": {"eliminate": 1e3 * sum(elim_s) / args.steps, "echelon": 1e3 * sum(ech_s) / args.steps},
        }
        if world == 1 and not args.no_cpu_baseline:
            out["cpu_baseline"] = cpu_baseline(steps, fma_top, args)
        print(json.dumps(out))
    for e in engines:
        e.close()
    if world > 1:
        import torch.distributed as dist
        dist.destroy_process_group()


def cpu_baseline(steps, fma_top, args):
    """The oracle port on a bounded sample (~args.ref_seconds of CPU work), rank 0, N=1 only."""
    from oracle import oracle
    oracle.build()
    i = min(range(len(steps)), key=lambda k: fma_top[k])
    s, f = steps[i], fma_top[i]
    every = max(1, int(f / 8e8 / args.ref_seconds))
    sample, frac = subsample_rows(s, every)
    nnz_sample = int(s.row_ptr[s.n_top]) * frac + (sample.nnz - int(s.row_ptr[s.n_top]))
    _, fma, sec = oracle.reduce(sample)
    return {"value": nnz_sample / sec, "unit": "nnz/s", "cores": 1, "kind": "port", "cpu": cpu_model(),
            "sample": f"{args.system} F4 step {s.n_rows}x{s.n_cols} ({s.nnz} nnz), all pivot rows + every {every}-th of the "
                      f"{s.n_bot} rows to reduce ({fma:.3g} multiply-adds, {sec:.1f} s); nnz counted = that fraction of nnz_in",
            "fma_per_s": fma / sec,
            "note": "oracle port; the reference's own engine is closed source. Published end-to-end (closed binary, "
                    "Ryzen 5 3600, 1 thread): cyclic9-15 17.8 s (README.md:18)"}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--system", default="cyclic9-15")
    ap.add_argument("--min-nnz", type=int, default=15_000_000)
    ap.add_argument("--max-matrices", type=int, default=0)
    ap.add_argument("--max-spairs", type=int, default=2000, help="pair cap of the driver run that produces the matrices (0 = none)")
    ap.add_argument("--ref-seconds", type=float, default=6.0)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-bcast", action="store_true", help="N>1: every rank uploads the step itself (no NCCL broadcast)")
    ap.add_argument("--opt", nargs="*", default=[], help="engine options key=value (gamba_cuda_set_option)")
    args = ap.parse_args()
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if args.impl == "reference":
        run_reference(args, rank, world)
    else:
        run_ours(args, rank, world, local_rank)


if __name__ == "__main__":
    main()
