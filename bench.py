#!/usr/bin/env python
"""Benchmark of the hot path: F4 linear-algebra throughput, nnz reduced per second (BASELINE.json's metric, quoted on
cyclic9 and kat15).

Workload (config.workload): EVERY F4 step matrix of cyclic9-15 (default options, 46 steps) and EVERY 3RD step matrix of
kat15-15 run with `--max-spairs 0` (same basis; 15 steps of up to 17 k rows to reduce against 500 k pivot rows instead of 60
steps of 2000 rows: the setting the GPU wants, SURVEY.md §8(d), and 43 s instead of 72 s end to end). The whole kat15-15 run is
4.6e9 nnz and 15 s of kernels per pass - too much for the K + W passes the driver asks for - so every 3rd step stands for it: an
unbiased sample from the small steps of the first degrees to the 517 k x 517 k ones of the last. Both are produced on the box by
the product driver (`gamba --dump-steps`, untimed workload generation).
One bench "step" = one pass of the engine over all those matrices in F4 order, one engine per system, so that every choice the
engine makes from the previous step (kernel choice, predicted rank of the Monte-Carlo echelon) is made as in a real run.

  value     Σ nnz_in / time with the index arrays of every step already resident in HBM (GAMBA_CUDA_STEP_DEVICE_INDICES;
            coefficient arrays resident through the engine's coeff_cache): staging kernels + operand builds + elimination +
            echelon + result D2H are all inside the timed region
  e2e       the same through gamba_cuda_reduce from pinned HOST buffers: H2D + everything above
  roofline  dominant kernel = limb_gemm_tc_kernel (every dense mod-p product of the elimination stage, tcgen05 kind::i8):
            int8 multiply-adds of all its launches / CUDA-event time of the stage without the operand builds / int8 tensor
            peak MEASURED in the same run (gamba_cuda_measure_int8_peak); the survey's algorithmic-bytes figure (§8(d)) beside it
            (on the steps that run on Monte-Carlo combinations of their rows the products are smaller by n_bot / K: the fraction
            says how well the launches that remain use the tensor cores, not how much work was avoided)
  cpu_baseline  the CPU oracle (a port: the reference's engine is closed source), single thread, on a bounded sample
  parity_check  untimed: every matrix reduced once more by an independent configuration on rank 0 (N > 1: a world_size = 1
            engine; N = 1: the sparse-kernel path, panel = 0) and compared by checksum

`--impl reference` runs the oracle's checker mode with all host threads on bounded samples of the SAME matrices (config is
identical; `cpu_baseline.sample` says what a step was).
"""
from __future__ import annotations

import argparse
import glob
import json
import os
import re
import shutil
import subprocess
import sys
import tempfile
import threading
import time
import zlib

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

import numpy as np  # noqa: E402

METRIC = "F4 lin-alg nnz reduced/s"
WORK_DIR = os.path.join(tempfile.gettempdir(), "gamba_b200_bench")
DEFAULT_SYSTEMS = "cyclic9-15:2000:1,kat15-15:0:3"          # system:max_spairs:every-nth-step
GAMBA = os.path.join(ROOT, "gamba_b200", "bin", "gamba")
README_SECONDS = {"cyclic9-15": 17.8, "cyclic9-31": 26.1, "kat15-15": 318.5, "kat15-31": 423.3, "eco15-31": 97.7,
                  "kat14-15": 59.4, "eco14-15": 16.6}   # closed reference binary, Ryzen 5 3600, 1 thread (README.md:18-23,42-47)


def measured_peaks():
    try:
        with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as f:
            return float(json.load(f)["hbm_gbs"]), "measured (MEASURED_PEAKS.json)"
    except Exception:  # noqa: BLE001
        return 6650.0, "fallback (B200_PROFILING.md)"


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


# ---------------------------------------------------------------------------------------------- workload

def all_cores():
    """subprocess preexec: the product driver is multi-threaded on the host; torch / OpenMP may have narrowed this process's
    CPU affinity, which a child would inherit."""
    try:
        os.sched_setaffinity(0, range(os.cpu_count() or 1))
    except (AttributeError, OSError):
        pass


def parse_systems(spec: str):
    out = []
    for item in spec.split(","):
        f = item.split(":")
        out.append((f[0], int(f[1]) if len(f) > 1 and f[1] else 2000, int(f[2]) if len(f) > 2 and f[2] else 1))
    return out


def generate_workload(system: str, max_spairs: int, min_nnz: int, rank: int, world: int, every: int = 1):
    """Step matrices of `system`, produced by the product driver on this box (rank 0; the others wait). Returns
    (step files in F4 order, seconds of the driver run that wrote them)."""
    from gamba_b200 import systems
    d = os.path.join(WORK_DIR, f"{system}_sp{max_spairs}_min{min_nnz}_every{every}")
    done = os.path.join(d, "DONE")
    if rank == 0 and not os.path.exists(done):
        shutil.rmtree(d, ignore_errors=True)
        os.makedirs(d)
        inp = systems.write_system(system, os.path.join(d, system + ".txt"))
        t = time.time()
        r = subprocess.run([GAMBA, "-i", inp, "-o", os.path.join(d, "out.txt"), "-v", "1", "--dump-steps", d,
                            "--dump-min-nnz", str(min_nnz), "--dump-every", str(every), "--max-spairs", str(max_spairs)],
                           capture_output=True, text=True, preexec_fn=all_cores)
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
    """[(system, max_spairs, [Step ...])], description. `synthetic:ROWSxCOLS[:density[:bits]]` gives BASELINE.json
    configs[4] (one F4-shaped random matrix, every rank generates the same one from the seed)."""
    from gamba_b200 import load_step
    from gamba_b200.stepfile import synthetic_step_large
    if args.systems.startswith("synthetic:"):
        parts = args.systems.split(":")
        nr, nc = (int(x) for x in parts[1].lower().split("x"))
        dens = float(parts[2]) if len(parts) > 2 else 0.005
        p = {"15": 32003, "31": 2147483647}[parts[3] if len(parts) > 3 else "31"]
        return ([("synthetic", 0, [synthetic_step_large(nr, nc, dens, p)])],
                f"synthetic F4-shaped matrix {nr} x {nc}, density {dens}, p={p}, seed 967557673")
    groups, desc = [], []
    for name, sp, every in parse_systems(args.systems):
        files = generate_workload(name, sp, args.min_nnz, rank, world, every)
        steps = [load_step(f) for f in files]
        if args.max_matrices:
            keep = set(sorted(range(len(steps)), key=lambda i: -steps[i].nnz)[: args.max_matrices])
            steps = [s for i, s in enumerate(steps) if i in keep]
        groups.append((name, sp, steps))
        desc.append(f"{name}" + ("" if sp == 2000 else f" --max-spairs {sp}") + f": {'all ' if not args.min_nnz and not args.max_matrices and every == 1 else ''}"
                    + (f"every {every}th step, " if every > 1 else "") +
                    f"{len(steps)} F4 step matrices" + (f" with >= {args.min_nnz} nnz" if args.min_nnz else "") +
                    f", {sum(s.nnz for s in steps)} nnz, largest {max(s.n_rows for s in steps)} x {max(s.n_cols for s in steps)}")
    return groups, "; ".join(desc)


def config_of(args, wl_desc):
    """Identical in both arms (the driver compares it)."""
    return {"workload": wl_desc, "systems": args.systems, "min_nnz": args.min_nnz, "max_matrices": args.max_matrices,
            "order": "one pass = every matrix once, in F4 order",
            "l2": "inputs larger than L2: a pass cycles through every matrix (GBs of index arrays and operands) before a matrix is seen again"}


def bytes_alg(step, fma_top: int) -> float:
    """fma_top*(4+w) + nnz_in*(4+w) + n_bot*n_nonpiv*8."""
    w = 2 if step.p < 65536 else 4
    return fma_top * (4 + w) + step.nnz * (4 + w) + step.n_bot * step.n_nonpiv * 8


def sample_nnz(step, sample) -> float:
    """nnz_in counted for a thinned step: the A|B part in proportion to the rows kept, plus the rows kept."""
    top = int(step.row_ptr[step.n_top])
    return top * (sample.n_bot / max(step.n_bot, 1)) + (sample.nnz - top)


def checksum(rows) -> int:
    c = zlib.crc32(np.ascontiguousarray(rows.row_ptr).view(np.uint8))
    c = zlib.crc32(np.ascontiguousarray(rows.col).view(np.uint8), c)
    return zlib.crc32(np.ascontiguousarray(rows.val).view(np.uint8), c)


# ---------------------------------------------------------------------------------------------- reference arm

def choose_thinning(groups, threads: int, target_s: float):
    """Rows kept per matrix so that one pass of the oracle over all matrices costs about target_s: calibrated on a
    1/64 sample of every matrix (fma_top of the sample / seconds), then scaled."""
    from gamba_b200.stepfile import thin_step
    from oracle import oracle
    est = 0.0
    for _, _, steps in groups:
        for s in steps:
            if s.n_bot < 64 or s.nnz < 200000:
                _, _, sec = oracle.reduce(s, threads=threads)
                est += sec
            else:
                _, _, sec = oracle.reduce(thin_step(s, 64), threads=threads)
                est += sec * 64
    return max(1, int(np.ceil(est / target_s)))


def run_reference(args, rank: int, world: int):
    """Reference arm: the CPU oracle (checker mode, all host threads) on bounded samples of the same matrices."""
    if rank != 0:          # under torchrun rank 0 alone runs the CPU arm; the other ranks exit without work
        return
    groups, wl_desc = load_workload(args, 0, 1)
    from gamba_b200.stepfile import thin_step
    from oracle import oracle
    oracle.build()
    threads = oracle.host_threads()
    every = choose_thinning(groups, threads, args.ref_seconds)
    samples, nnz_s = [], 0.0
    for _, _, steps in groups:
        for s in steps:
            t = thin_step(s, every) if s.n_bot >= 2 * every else s
            samples.append(t); nnz_s += sample_nnz(s, t)
    for _ in range(min(args.warmup, 1)):
        for t in samples:
            oracle.reduce(t, threads=threads)
    t0 = time.perf_counter()
    for _ in range(args.steps):
        for t in samples:
            oracle.reduce(t, threads=threads)
    wall = time.perf_counter() - t0
    value = nnz_s * args.steps / wall
    desc = (f"every matrix of the workload with all its pivot rows and every {every}-th of its rows to reduce "
            f"({len(samples)} matrices per step); nnz counted = nnz(A|B) in that proportion + nnz of the rows kept")
    print(json.dumps({
        "impl": "reference", "metric": METRIC, "value": value, "unit": "nnz/s", "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": 1e3 * wall / args.steps, "higher_is_better": True, "scaling": "strong",
        "vs_baseline": None, "dtype": "u64 (CPU oracle accumulators) over F_p", "data": "synthetic (generated systems, no dataset)",
        "config": config_of(args, wl_desc),
        "cpu_baseline": {"value": value, "unit": "nnz/s", "cores": threads, "kind": "port", "sample": desc, "cpu": cpu_model(),
                         "note": "oracle port in checker mode (all host threads); the reference's own engine (linalg_v3.hpp, "
                                 "kernel/*) is closed source and single-threaded"},
        "e2e": {"value": value, "unit": "nnz/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }), flush=True)


# ---------------------------------------------------------------------------------------------- our arm

def e2e_seconds(groups):
    """Wall seconds of the product binary end to end per system (README.md rows beside them), rank 0, N = 1."""
    from gamba_b200 import systems
    out = {}
    for name, sp, _ in groups:
        if name == "synthetic":
            continue
        d = tempfile.mkdtemp(prefix="gamba_e2e_")
        inp = systems.write_system(name, os.path.join(d, name + ".txt"))
        t = time.perf_counter()
        r = subprocess.run([GAMBA, "-i", inp, "-o", os.path.join(d, "out.txt"), "-v", "1", "--max-spairs", str(sp)],
                           capture_output=True, text=True, preexec_fn=all_cores)
        wall = time.perf_counter() - t
        m = re.search(r"linalg \(wall\)\s+([0-9.]+) sec\s+\[device-only ([0-9.]+)\]", r.stdout)
        extra = {}
        if name.startswith(("kat15", "eco15")) and sp != 0:      # the GPU-preferred setting: fewer, larger steps, same basis
            t = time.perf_counter()
            r0 = subprocess.run([GAMBA, "-i", inp, "-o", os.path.join(d, "out0.txt"), "-v", "0", "--max-spairs", "0"], capture_output=True, text=True,
                                preexec_fn=all_cores)
            same = r0.returncode == 0 and open(os.path.join(d, "out0.txt"), "rb").read() == open(os.path.join(d, "out.txt"), "rb").read()
            extra = {"seconds_max_spairs_0": time.perf_counter() - t if r0.returncode == 0 else None, "same_basis_as_default_run": same}
        out[name] = {"seconds": wall if r.returncode == 0 else None, "max_spairs": sp, **extra,
                     "linalg_wall_s": float(m.group(1)) if m else None, "linalg_device_s": float(m.group(2)) if m else None,
                     "reference_binary_seconds": README_SECONDS.get(name),
                     "reference_binary_note": "closed GamBa binary, Ryzen 5 3600, 1 thread, default options (README.md:18-23,42-47)"}
        shutil.rmtree(d, ignore_errors=True)
    return out


def log(msg: str):
    if os.environ.get("GAMBA_BENCH_VERBOSE"):
        print(f"[bench rank {os.environ.get('RANK', '0')}] {msg}", file=sys.stderr, flush=True)


def run_ours(args, rank: int, world: int, local_rank: int):
    import torch
    from gamba_b200 import LinalgEngine, _lib, parallel
    from gamba_b200.stepfile import Step
    if not torch.cuda.is_available():
        raise RuntimeError("bench.py needs a CUDA device: the product has no CPU path")
    torch.cuda.set_device(local_rank)
    if world > 1:
        import torch.distributed as dist
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    # end-to-end seconds of the product binary first, while this process holds neither the workload nor a CUDA context
    # (measured after the timed legs the same runs took twice as long: the host phases of the driver are memory-bound)
    e2e_s = None
    if world == 1 and rank == 0 and not args.no_e2e_seconds and not args.systems.startswith("synthetic:"):
        e2e_s = e2e_seconds([(n, sp, None) for n, sp, _ in parse_systems(args.systems)])
    groups, wl_desc = load_workload(args, rank, world)
    total_nnz = sum(s.nnz for _, _, steps in groups for s in steps)
    n_mat = sum(len(steps) for _, _, steps in groups)

    def sync():
        torch.cuda.synchronize()
        if world > 1:
            import torch.distributed as dist
            dist.barrier()

    gather = parallel.make_allgather() if world > 1 else None
    bcast = parallel.make_broadcast() if world > 1 and not args.no_bcast else None

    def make_engine(p, **kw):
        e = LinalgEngine(p, device=local_rank, **kw)
        e.set_option("coeff_cache", 1)          # as the product driver does: basis coefficient arrays stay in HBM
        for kv in args.opt:
            k, v = kv.split("=")
            e.set_option(k, int(v))
        return e

    def pinned(s: Step) -> Step:                # host buffers the e2e leg copies from
        keep, out = [], []
        for a in (s.row_ptr, s.col, s.coeff_id, s.coeff_len, s.pool):
            t = torch.from_numpy(np.ascontiguousarray(a)).pin_memory()
            out.append(t.numpy()); keep.append(t)
        st = Step(s.p, s.n_top, s.n_bot, s.n_cols, *out)
        st._pins = keep
        return st

    engines, residents, hosts = [], [], []
    for _, _, steps in groups:
        e = make_engine(steps[0].p, rank=rank, world_size=world)
        if gather:
            e.set_allgather(gather)
        engines.append(e)
        hs = [pinned(s) for s in steps]
        hosts.append(hs)
        residents.append([e.make_resident(s, local_rank) for s in hs])

    def launches():
        return sum(e.counters()["kernel_launches"] for e in engines)

    log(f"workload loaded: {n_mat} matrices, {total_nnz} nnz")
    # ---- parity check (untimed): an independent configuration on rank 0 must give the same rows ----------------------
    parity = {"status": "skipped"}
    if not args.no_parity:
        sums = []
        for e, rs in zip(engines, residents):
            sums.append([checksum(e.reduce_resident(r)) for r in rs])
        if rank == 0:
            bad, n_cmp = [], 0
            for gi, (name, _, steps) in enumerate(groups):
                chk = make_engine(steps[0].p)                 # world_size = 1
                if world == 1:
                    chk.set_option("panel", 0)                # N = 1: the sparse-kernel path instead of the tensor-core solve
                for si, s in enumerate(hosts[gi]):
                    if world == 1 and s.nnz > args.parity_max_nnz:
                        continue
                    n_cmp += 1
                    if checksum(chk.reduce(s)) != sums[gi][si]:
                        bad.append(f"{name}[{si}]")
                chk.close()
            parity = {"status": "ok" if not bad else "MISMATCH " + ",".join(bad), "matrices": n_cmp,
                      "against": ("world_size=1 reduction of the same matrices on rank 0" if world > 1 else
                                  f"sparse-kernel path (panel=0) on the matrices with <= {args.parity_max_nnz} nnz")}
        sync()

    log(f"parity check: {parity}")
    # ---- value: index arrays resident in HBM -------------------------------------------------------------------------
    def run_pass(collect=None):
        for gi, (e, rs) in enumerate(zip(engines, residents)):
            for si, r in enumerate(rs):
                e.reduce_resident(r, copy=False)     # the result stays in the engine-owned host buffers the C ABI hands out
                if collect is not None:
                    collect(gi, si, e)

    for _ in range(args.warmup):
        run_pass()
    sync()
    per = {}

    def collect(gi, si, e):
        c = e.counters()
        per[(gi, si)] = (e.last["seconds_eliminate"], e.last["seconds_echelon"], e.last["seconds_device"], c["fma_top"],
                         c["dense_macs"], c["seconds_operands_last"], c["solve_mode"], c["seconds_sparse_last"], c["mc_combinations"])
    l0 = launches()
    with ClockSampler(local_rank) as clk:
        sync()
        t0 = time.perf_counter()
        for i in range(args.steps):
            run_pass(collect if i == args.steps - 1 else None)
        sync()
        t_res = time.perf_counter() - t0
    n_launch = launches() - l0
    clocks = clk.summary()

    log(f"value leg done: {t_res:.2f} s for {args.steps} passes")
    # ---- e2e: pinned host buffers in, host buffers out ---------------------------------------------------------------
    if bcast:
        for e in engines:
            e.set_broadcast(bcast)
    e2e_steps = max(1, min(args.steps, args.e2e_steps))

    def run_pass_host():
        h2d = d2h = 0
        for e, hs in zip(engines, hosts):
            for s in hs:
                e.reduce(s if (rank == 0 or not bcast) else None, copy=False)
                h2d += e.last["h2d_bytes"]; d2h += e.last["d2h_bytes"]
        return h2d, d2h
    run_pass_host()
    sync()
    t0 = time.perf_counter()
    h2d = d2h = 0
    for _ in range(e2e_steps):
        a, b = run_pass_host()
        h2d += a; d2h += b
    sync()
    t_e2e = time.perf_counter() - t0

    log(f"e2e leg done: {t_e2e:.2f} s for {e2e_steps} passes")
    # the engines (tens of GB of operands) and the resident steps go before anything else wants the GPU (e2e_seconds runs the product binary)
    for e in engines:
        e.close()
    residents.clear()
    hosts.clear()                  # page-locked copies of the workload
    torch.cuda.empty_cache()
    keys = sorted(per)
    vals = np.array([per[k] for k in keys], dtype=np.float64)
    if rank == 0 and os.environ.get("GAMBA_BENCH_VERBOSE"):
        for (gi, si), v in zip(keys, vals):
            st = groups[gi][2][si]
            if st.nnz >= 20_000_000:
                log(f"{groups[gi][0]}[{si}] {st.n_rows}x{st.n_cols} n_bot={st.n_bot}: mode {int(v[6])} eliminate {1e3 * v[0]:.1f} ms (operands {1e3 * v[5]:.1f}) echelon {1e3 * v[1]:.1f} ms")
    if world > 1:
        import torch.distributed as dist
        tt = torch.tensor([t_res, t_e2e], device="cuda", dtype=torch.float64)
        dist.all_reduce(tt, op=dist.ReduceOp.MAX)
        t_res, t_e2e = float(tt[0]), float(tt[1])
        vt = torch.tensor(vals, device="cuda", dtype=torch.float64)
        vmax = vt.clone(); dist.all_reduce(vmax, op=dist.ReduceOp.MAX)
        vsum = vt.clone(); dist.all_reduce(vsum, op=dist.ReduceOp.SUM)
        vals = vmax.cpu().numpy()
        vals[:, 3] = vsum.cpu().numpy()[:, 3]          # fma_top: every rank counted its own rows
        vals[:, 4] = vsum.cpu().numpy()[:, 4]          # dense multiply-adds: summed over the ranks

    if rank == 0:
        hbm_peak, hbm_src = measured_peaks()
        lib = _lib.load()
        import ctypes as C
        steps_flat = [groups[gi][2][si] for gi, si in keys]
        nl2 = np.array([4.0 if s.p < 65536 else 16.0 for s in steps_flat])
        peak_n = 128 if steps_flat[0].p < 65536 else 64
        pk = C.c_double(0.0)
        if lib.gamba_cuda_measure_int8_peak(local_rank, peak_n, C.byref(pk)):
            raise RuntimeError("gamba_cuda_measure_int8_peak: " + lib.gamba_cuda_last_error().decode())
        int8_peak = pk.value                                         # TMAC/s
        elim, ech, dev, fma_top, dense, opnd, mode, sparse, mc_k = (vals[:, i] for i in range(9))
        gemm_s = float(np.sum(np.where(mode == 2, elim - opnd, np.maximum(elim - sparse - opnd, 0.0))))   # the products' share of the stage
        int8_macs = float(np.sum(dense * nl2))
        achieved_tops = 2.0 * int8_macs / max(gemm_s, 1e-12) / 1e12 / world          # per GPU
        alg = float(sum(bytes_alg(s, f) for s, f in zip(steps_flat, fma_top)))
        out = {
            "metric": METRIC, "value": total_nnz * args.steps / t_res, "unit": "nnz/s", "n_gpus": world,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * t_res / args.steps,
            "higher_is_better": True, "scaling": "strong", "vs_baseline": None,
            "dtype": "u32 residues mod p; u8 limbs -> s32 accumulators on the tensor cores (tcgen05 kind::i8) in every dense product "
                     "and the echelon; 32-bit delayed-reduction lanes in the sparse kernels",
            "data": "synthetic (generated systems, no dataset)",
            "config": config_of(args, wl_desc),
            "engine": {"parallelism": (f"C|D rows sharded r % {world}; value: every rank holds the step; e2e: rank 0 uploads, NCCL broadcast; "
                                       f"NCCL all-gather of D' rows, echelon replicated") if world > 1 else "1 GPU",
                       "matrices_per_step": n_mat, "nnz_per_step": total_nnz,
                       "solve_modes": {"sparse": int(np.sum(mode == 0)), "sparse_blocked": int(np.sum(mode == 1)),
                                       "tensor_core_blocked_solve": int(np.sum(mode == 2))},
                       "monte_carlo": {"matrices": int(np.sum(mc_k > 0)), "combinations": int(np.sum(mc_k)),
                                       "rows_they_replace": int(sum(s.n_bot for s, k in zip(steps_flat, mc_k) if k > 0)),
                                       "note": "steps whose rows to reduce were replaced by K = predicted rank + margin random combinations "
                                               "(kernels_mc.cuh; accepted only with K - rank >= 32, result bit-exact)"}},
            "e2e": {"value": total_nnz * e2e_steps / t_e2e, "unit": "nnz/s", "h2d_bytes_per_step": h2d // e2e_steps,
                    "d2h_bytes_per_step": d2h // e2e_steps, "ms_per_step": 1e3 * t_e2e / e2e_steps, "steps": e2e_steps},
            "gpu_launches": n_launch,
            "clocks": clocks,
            "parity_check": parity,
            "roofline": {"kernel": "limb_gemm_tc_kernel: every dense mod-p product of the elimination stage (blocked triangular solve: "
                                   "earlier-tile products, inverse-tile products, recursive doubling; D' = D + M B), tcgen05 kind::i8 on u8 limb planes",
                         "bound": "tensor", "achieved": achieved_tops, "peak": 2.0 * int8_peak, "unit": "TOP/s (int8, 2 per multiply-add)",
                         "frac": achieved_tops / (2.0 * int8_peak),
                         "peak_source": f"measured in this run: tcgen05.mma kind::i8 128x{peak_n}x32 from shared memory on all SMs "
                                        f"(gamba_cuda_measure_int8_peak); nominal dense int8 4500 TOP/s",
                         "traffic": None, "traffic_note": "not measured inside bench.py; ncu captures of the kernel are under profiles/",
                         "launch_time_basis": "CUDA events on the engine's stream around the stage, minus the operand builds; includes the small "
                                              "scatter / base-inverse kernels between the products (conservative)",
                         "int8_macs_per_step": int8_macs, "avg_stage_ms_per_matrix": 1e3 * gemm_s / max(len(keys), 1),
                         "share_of_device_time": float(np.sum(elim) / max(np.sum(dev), 1e-12)),
                         "hbm_alg": {"note": "SURVEY.md §8(d) figure: algorithmic bytes of the sparse row-by-row elimination / elimination time / "
                                             "measured HBM peak; > 1 because the stage is dense products, not streamed pivot rows",
                                     "achieved": alg / max(float(np.sum(elim)), 1e-12) / 1e9 / world, "peak": hbm_peak, "unit": "GB/s",
                                     "frac": alg / max(float(np.sum(elim)), 1e-12) / 1e9 / world / hbm_peak, "peak_source": hbm_src,
                                     "fma_top_per_step": float(np.sum(fma_top))}},
            "breakdown_ms_per_step": {"staging_kernels": 1e3 * float(np.sum(dev - elim - ech)), "operand_builds": 1e3 * float(np.sum(opnd)),
                                      "eliminate_without_operands": 1e3 * float(np.sum(elim - opnd)), "echelon": 1e3 * float(np.sum(ech)),
                                      "device_total": 1e3 * float(np.sum(dev))},
        }
        if e2e_s is not None:
            out["e2e_seconds"] = e2e_s
        if world == 1 and not args.no_cpu_baseline:
            out["cpu_baseline"] = cpu_baseline(groups, args)
        print(json.dumps(out), flush=True)
    if world > 1:
        import torch.distributed as dist
        dist.destroy_process_group()


def cpu_baseline(groups, args):
    """The oracle port, single thread (the reference is single-threaded, README.md:9,71), on a bounded sample, rank 0, N=1."""
    from gamba_b200.stepfile import thin_step
    from oracle import oracle
    oracle.build()
    name, _, steps = groups[0]
    # the matrices of the first system with the most work (nnz x rows to reduce), whole, until about cpu_seconds of oracle time are
    # reached; if the heaviest alone costs more it is thinned (all pivot rows stay: the cost of a row to reduce does not depend on
    # how many others there are). Costs are estimated on every 64th row first.
    work = lambda x: float(x.nnz) * max(x.n_bot, 1)
    order = sorted(steps, key=lambda x: -work(x))
    s0 = order[0]
    est0 = oracle.reduce(thin_step(s0, 64))[2] * 64 if s0.n_bot >= 128 else 0.0     # an upper estimate: the set-up cost counts 64 times
    chosen = []          # (matrix, sample, every, multiply-adds, seconds)
    if est0 > 2.0 * args.cpu_seconds:
        every = min(64, int(np.ceil(est0 / args.cpu_seconds)))
        sample = thin_step(s0, every)
        _, fma, sec = oracle.reduce(sample)
        chosen.append((s0, sample, every, fma, sec))
    else:
        _, fma, sec = oracle.reduce(s0)
        chosen.append((s0, s0, 1, fma, sec))
        total = sec
        for s in order[1:16]:
            if total >= args.cpu_seconds:
                break
            if s.n_bot >= 128 and total + oracle.reduce(thin_step(s, 64))[2] * 64 > 2.5 * args.cpu_seconds:
                continue         # (upper estimate) would take the sample beyond its bound: a lighter matrix instead
            _, f2, t2 = oracle.reduce(s)
            chosen.append((s, s, 1, f2, t2)); total += t2
    nnz_s = sum(sample_nnz(s, sample) for s, sample, _, _, _ in chosen)
    fma_s = float(sum(c[3] for c in chosen)); sec_s = float(sum(c[4] for c in chosen))
    every0 = chosen[0][2]
    what = (f"all pivot rows + every {every0}-th of the {s0.n_bot} rows to reduce of the F4 step {s0.n_rows}x{s0.n_cols} ({s0.nnz} nnz)" if every0 > 1 else
            f"the {len(chosen)} F4 step matrices with the most work, whole (largest {s0.n_rows}x{s0.n_cols}, {s0.nnz} nnz)")
    return {"value": nnz_s / sec_s, "unit": "nnz/s", "cores": 1, "kind": "port", "cpu": cpu_model(),
            "sample": f"{name}: {what} ({fma_s:.3g} multiply-adds, {sec_s:.1f} s); nnz counted = nnz(A|B) in the proportion of the rows kept + nnz of the rows kept",
            "fma_per_s": fma_s / sec_s,
            "note": "oracle port, timing mode; the reference's own engine is closed source"}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=3)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--systems", default=DEFAULT_SYSTEMS, help="system:max_spairs:every[,...] or synthetic:ROWSxCOLS[:density[:bits]]")
    ap.add_argument("--min-nnz", type=int, default=0, help="only step matrices with at least this many non-zeros (default: all)")
    ap.add_argument("--max-matrices", type=int, default=0, help="only the N largest matrices of each system (default: all)")
    ap.add_argument("--e2e-steps", type=int, default=3, help="passes of the host-buffer leg (at most --steps)")
    ap.add_argument("--ref-seconds", type=float, default=6.0, help="--impl reference: CPU seconds one sampled pass should cost")
    ap.add_argument("--cpu-seconds", type=float, default=12.0, help="cpu_baseline: CPU seconds of the sample")
    ap.add_argument("--parity-max-nnz", type=int, default=60_000_000, help="N=1 parity check: matrices above this are skipped (sparse path: seconds each)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-e2e-seconds", action="store_true")
    ap.add_argument("--no-parity", action="store_true")
    ap.add_argument("--no-bcast", action="store_true", help="N>1 e2e leg: every rank uploads the step itself (no NCCL broadcast)")
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
