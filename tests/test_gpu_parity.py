"""Parity tests proper: the CUDA path, called through the C ABI, against the CPU oracle — bit-exact
(integer arithmetic mod p; no tolerance). Step matrices come from the host driver run on the reference's
systems; the oracle rows are what the oracle driver wrote beside them."""
import glob
import os

import numpy as np
import pytest

from gamba_b200.stepfile import Rows, Step, load_rows, load_step, synthetic_step

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def engines():
    from gamba_b200 import LinalgEngine
    cache = {}

    def get(p, **kw):
        key = (p,) + tuple(sorted(kw.items()))
        if key not in cache:
            cache[key] = LinalgEngine(p, **kw)
        return cache[key]
    yield get
    for e in cache.values():
        e.close()


def check_dir(d, eng, min_nnz=0, final=True):
    n = 0
    for f in sorted(glob.glob(os.path.join(d, "step_*.bin"))):
        is_final = f.endswith("step_final.bin")
        if is_final and not final:
            continue
        s = load_step(f)
        if s.nnz < min_nnz:
            continue
        want = load_rows(f.replace("step_", "rows_"))
        got = eng.reduce(s, final=is_final)
        assert got.same_as(want), f
        assert eng.last["zero_reductions"] == s.n_bot - want.n_new
        n += 1
    return n


# 15-, 31-, 16-, 8- and 3-bit primes: the reference's u16 / u32 arithmetic paths (CMakeLists.txt:81-90)
@pytest.mark.parametrize("name", ["cyclic6-15", "cyclic7-15", "cyclic7-31", "cyclic7-8", "cyclic7-3", "kat9-16",
                                  "kat9-31", "eco10-31", "noon6-31", "reimer6-31", "henrion6-15"])
def test_every_step_matches_oracle(step_dirs, engines, name):
    d = step_dirs(name)
    s0 = load_step(sorted(glob.glob(os.path.join(d, "step_*.bin")))[0])
    assert check_dir(d, engines(s0.p)) > 0


@pytest.mark.parametrize("name", ["cyclic8-15", "cyclic8-31", "kat10-15", "eco11-16"])
def test_larger_systems_steps(step_dirs, engines, name):
    d = step_dirs(name, ("--dump-min-nnz", "100000"))
    s0 = load_step(sorted(glob.glob(os.path.join(d, "step_*.bin")))[0])
    assert check_dir(d, engines(s0.p)) >= 5


def test_max_spairs_zero_mode(step_dirs, engines):
    """The reference's second test mode (test/test_output.sh:24): bigger, fewer steps."""
    d = step_dirs("cyclic7-15", ("--max-spairs", "0"))
    assert check_dir(d, engines(32003)) > 0


@pytest.mark.parametrize("tile", [128, 256, 1024])
def test_forced_small_tiles_exercise_deferred_passes(step_dirs, tile):
    """Column tiles far narrower than the matrix: every pivot row is applied in pieces across many tiles."""
    from gamba_b200 import LinalgEngine
    eng = LinalgEngine(32003)
    eng.set_option("tile_cols", tile)
    d = step_dirs("cyclic7-15")
    assert check_dir(d, eng) > 0
    assert eng.counters()["n_tiles"] >= 2
    eng.close()


@pytest.mark.parametrize("warps", [2, 4, 32])
@pytest.mark.parametrize("name", ["cyclic7-15", "cyclic7-31"])
def test_warps_per_row_of_the_sparse_kernel(step_dirs, name, warps):
    """elim_kernel with 2 / 4 / 32 warps sharing a row (default 8; 32 is what steps with fewer rows than SMs against >= 65536
    pivot rows get: one CTA per row, and a warp retires a chunk per ~450 cycles whatever is in flight), with and without the
    tensor-core product for the B half."""
    from gamba_b200 import LinalgEngine
    d = step_dirs(name)
    p = load_step(sorted(glob.glob(os.path.join(d, "step_*.bin")))[0]).p
    for schur_min_top in (1, 1 << 30):
        eng = LinalgEngine(p)
        eng.set_option("elim_warps", warps); eng.set_option("schur_min_top", schur_min_top); eng.set_option("panel", 0); eng.set_option("block", 0)
        assert check_dir(d, eng) > 0
        assert eng.counters()["solve_mode"] == 0
        eng.close()


def test_elimination_order_and_blockelim_steps(step_dirs, engines):
    d = step_dirs("cyclic5-15", ("-e", "2"))
    assert check_dir(d, engines(32003)) > 0


@pytest.mark.parametrize("p,cb", [(251, 1), (32003, 2), (65521, 2), (2147483647, 4), (7, 1)])
def test_coefficient_widths_and_montgomery_form(built, p, cb):
    """coeff_bytes 1/2/4 and the reference's Montgomery representation (field.hpp:41-51) at the boundary:
    R = 2^32 for 1-/2-byte coefficients, 2^64 for 4-byte ones; results come back in the same form."""
    from gamba_b200 import LinalgEngine, _lib
    from oracle import oracle
    s = synthetic_step(400, 520, 0.06, p, seed=5, top_frac=0.85, share=7)
    want, _, _ = oracle.reduce(s)
    e = LinalgEngine(p, coeff_bytes=cb)
    assert e.reduce(s).same_as(want)
    e.close()
    R = pow(2, 64 if cb == 4 else 32, p)
    sm = Step(s.p, s.n_top, s.n_bot, s.n_cols, s.row_ptr, s.col, s.coeff_id, s.coeff_len,
              (s.pool.astype(object) * R % p).astype(np.uint32))
    e = LinalgEngine(p, coeff_bytes=cb, repr=_lib.REPR_MONTGOMERY)
    got = e.reduce(sm)
    e.close()
    assert got.n_new == want.n_new and np.array_equal(got.col, want.col)
    assert np.array_equal(got.val, (want.val.astype(object) * R % p).astype(np.uint32))


def test_empty_and_ragged_inputs(built, engines):
    from oracle import oracle
    p = 32003
    eng = engines(p)
    s = synthetic_step(60, 90, 0.1, p, seed=2, top_frac=0.7, share=3)
    # no rows to reduce
    nt = s.n_top
    s0 = Step(p, nt, 0, s.n_cols, s.row_ptr[: nt + 1].copy(), s.col[: int(s.row_ptr[nt])].copy(), s.coeff_id[:nt].copy(),
              s.coeff_len, s.pool)
    r = eng.reduce(s0)
    assert r.n_new == 0 and list(r.row_ptr) == [0]
    # no pivot rows at all: plain echelon of a sparse block
    bot = slice(nt, s.n_rows)
    rp = (s.row_ptr[nt:] - s.row_ptr[nt]).astype(np.uint64)
    cols = s.col[int(s.row_ptr[nt]):].copy()
    keep = cols >= nt  # only non-pivot entries survive; rows get ragged (some may become short)
    lens = np.add.reduceat(keep.astype(np.int64), rp[:-1].astype(np.int64)) if len(rp) > 1 else np.zeros(0, np.int64)
    # coefficient arrays must follow the kept positions: materialise per-row arrays
    off = s.coeff_offsets()
    pool, clen, cid = [], [], []
    for i, r_ in enumerate(range(nt, s.n_rows)):
        a, b = int(s.row_ptr[r_]), int(s.row_ptr[r_ + 1])
        cf = s.pool[int(off[s.coeff_id[r_]]): int(off[s.coeff_id[r_]]) + (b - a)]
        k = keep[a - int(s.row_ptr[nt]): b - int(s.row_ptr[nt])]
        cfk = cf[k].copy()
        if len(cfk):
            cfk = (cfk.astype(np.int64) * pow(int(cfk[0]), -1, p) % p).astype(np.uint32)
        pool.append(cfk); clen.append(len(cfk)); cid.append(i)
    rp2 = np.zeros(len(clen) + 1, dtype=np.uint64); np.cumsum(clen, out=rp2[1:])
    s1 = Step(p, 0, len(clen), s.n_cols - nt, rp2, (cols[keep] - nt).astype(np.uint32), np.array(cid, np.uint32),
              np.array(clen, np.uint32), np.concatenate(pool) if pool else np.zeros(0, np.uint32))
    want, _, _ = oracle.reduce(s1)
    assert eng.reduce(s1).same_as(want)
    assert lens.sum() == s1.nnz


def test_rows_that_reduce_to_zero_and_duplicates(built, engines):
    """Rows to reduce that are copies / multiples-in-disguise of pivot rows vanish; duplicates give one pivot."""
    from oracle import oracle
    p = 2147483647
    s = synthetic_step(200, 300, 0.08, p, seed=9, top_frac=0.8, share=4)
    nt = s.n_top
    base = int(s.row_ptr[nt])
    rows = [(int(s.row_ptr[r]), int(s.row_ptr[r + 1]), int(s.coeff_id[r])) for r in [3, 7, nt, nt, nt + 1, 3]]
    rp = [base]
    col = [s.col[:base]]
    for a, b, _ in rows:
        col.append(s.col[a:b]); rp.append(rp[-1] + (b - a))
    s2 = Step(p, nt, len(rows), s.n_cols, np.concatenate([s.row_ptr[:nt], np.array(rp, np.uint64)]).astype(np.uint64),
              np.concatenate(col), np.concatenate([s.coeff_id[:nt], np.array([c for _, _, c in rows], np.uint32)]),
              s.coeff_len, s.pool)
    want, _, _ = oracle.reduce(s2)
    got = engines(p).reduce(s2)
    assert got.same_as(want) and got.n_new <= 2


def test_staged_reduce_is_repeatable(built, engines):
    from oracle import oracle
    s = synthetic_step(500, 700, 0.05, 32003, seed=4, top_frac=0.9, share=10)
    want, _, _ = oracle.reduce(s)
    eng = engines(32003)
    eng.stage(s)
    for _ in range(3):
        assert eng.reduce_staged().same_as(want)
    c = eng.counters()
    assert c["kernel_launches"] > 0 and c["fma_applied"] > 0


def test_fma_top_invariant_matches_oracle(built, engines):
    """fma_top (the roofline's unit of work) is a matrix invariant: the kernel's count
    of pivot hits x pivot-row length equals the oracle's."""
    from oracle import oracle
    for p, seed in [(32003, 6), (2147483647, 7)]:
        s = synthetic_step(700, 900, 0.04, p, seed=seed, top_frac=0.9, share=10)
        want, fma_top, _ = oracle.reduce(s)
        eng = engines(p)
        assert eng.reduce(s).same_as(want)
        assert eng.counters()["fma_top"] == fma_top


@pytest.mark.parametrize("p", [32003, 2147483647])
def test_full_size_properties(built, p):
    """BASELINE configs[4]-shaped synthetic step, too big for the oracle in seconds: size-independent
    properties instead — canonical form (monic, ascending pivots, pivot columns cleared) and idempotence:
    feeding the result back as rows to reduce (no pivot rows) returns it unchanged."""
    from gamba_b200 import LinalgEngine
    s = synthetic_step(6000, 9000, 0.01, p, seed=21, top_frac=0.92, share=25)
    eng = LinalgEngine(p)
    r = eng.reduce(s)
    assert 0 < r.n_new <= s.n_bot
    piv = r.col[r.row_ptr[:-1].astype(np.int64)]
    assert np.all(np.diff(piv.astype(np.int64)) > 0)
    assert np.all(r.val[r.row_ptr[:-1].astype(np.int64)] == 1)
    assert np.all(r.val > 0) and np.all(r.val < p)
    is_piv = np.zeros(s.n_nonpiv, bool); is_piv[piv] = True
    heads = np.zeros(len(r.col), bool); heads[r.row_ptr[:-1].astype(np.int64)] = True
    assert not np.any(is_piv[r.col] & ~heads)
    lens = np.diff(r.row_ptr.astype(np.int64)).astype(np.uint32)
    s2 = Step(p, 0, r.n_new, s.n_nonpiv, r.row_ptr, r.col, np.arange(r.n_new, dtype=np.uint32), lens, r.val)
    assert eng.reduce(s2).same_as(r)
    eng.close()


def mc_engine(p, mode):
    """Monte-Carlo combinations need one of the blocked solves: forced here on steps far below the sizes where the cost model picks them."""
    from gamba_b200 import LinalgEngine
    eng = LinalgEngine(p)
    eng.set_option("schur_min_top", 1)
    eng.set_option("mc", 2)
    if mode == "panel":
        eng.set_option("panel", 2); eng.set_option("panel_t", 128)
    else:
        eng.set_option("panel", 0); eng.set_option("block", 2)
    return eng


@pytest.mark.parametrize("mode", ["panel", "block"])
@pytest.mark.parametrize("name", ["cyclic8-15", "cyclic8-31", "kat10-15"])
def test_monte_carlo_combinations_are_bit_exact(step_dirs, name, mode):
    """Random combinations S (C|D) of the rows to reduce are eliminated instead of the rows themselves (README.md:5 of the
    reference: "Monte-Carlo F4"; kernels_mc.cuh), accepted only with a rank margin of 32, so the canonical RREF is the one of
    the exact computation. rank_hint = the true rank: one draw is enough; rank_hint far too small: the margin test must add
    combinations (top-up behind the first draw) or send the step back to the rows themselves."""
    d = step_dirs(name, ("--dump-min-nnz", "100000"))
    files = [f for f in sorted(glob.glob(os.path.join(d, "step_*.bin"))) if not f.endswith("step_final.bin")]
    eng = mc_engine(load_step(files[0]).p, mode)
    used = topped_up = exact = 0
    for f in files:
        s = load_step(f)
        want = load_rows(f.replace("step_", "rows_"))
        if s.n_bot < 256:
            continue
        for hint in (want.n_new, max(8, want.n_new // 3), 8):
            eng.set_option("rank_hint", hint)
            assert eng.reduce(s).same_as(want), (f, hint)
            assert eng.last["zero_reductions"] == s.n_bot - want.n_new
            c = eng.counters()
            assert c["solve_mode"] == (2 if mode == "panel" else 1)
            used += c["mc_combinations"] > 0 and c["mc_attempts"] == 1
            topped_up += c["mc_combinations"] > 0 and c["mc_attempts"] > 1
            exact += c["mc_combinations"] == 0
            if c["mc_combinations"]:
                assert c["mc_combinations"] >= want.n_new + 32 and c["mc_combinations"] * 2 <= s.n_bot + 64
    assert used > 0 and topped_up > 0, (used, topped_up, exact)
    if name.startswith("cyclic8"):       # step 18: 342 rows of rank 80 - from rank_hint 8 the draws reach half the rows: exact
        assert exact > 0
    eng.close()


def test_monte_carlo_off_and_final_steps_are_exact(step_dirs):
    """mc = 0 never draws; the final inter-reduction (rank = number of rows) is never drawn either."""
    d = step_dirs("cyclic8-15", ("--dump-min-nnz", "100000"))
    eng = mc_engine(32003, "panel")
    assert check_dir(d, eng) >= 5                      # includes step_final.bin
    eng.set_option("mc", 0)
    for f in sorted(glob.glob(os.path.join(d, "step_*.bin")))[:6]:
        if f.endswith("step_final.bin"):
            continue
        s = load_step(f)
        eng.set_option("rank_hint", 8)
        assert eng.reduce(s).same_as(load_rows(f.replace("step_", "rows_")))
        assert eng.counters()["mc_combinations"] == 0
    eng.close()


def test_31bit_prime_with_several_non_pivot_tiles(step_dirs, engines):
    """Regression (cyclic9-31 steps 9-10): two words per accumulator column for p >= 2^16 and more than
    one non-pivot tile — the tile was zeroed for the next pass while other threads still wrote it out."""
    d = step_dirs("cyclic9-31", ("--stop-after", "10", "--dump-min-nnz", "1000000"))
    eng = engines(2147483647)
    assert check_dir(d, eng) >= 2
    assert eng.counters()["n_tiles"] >= 3


@pytest.mark.parametrize("p", [32003, 2147483647])
def test_indirect_hand_over_form(built, engines, p):
    """ABI v2: rows stored anywhere (row_src), entries as caller handles (col_map) — the form the driver uses
    so that it never rewrites O(nnz) data (reference: rows are monomial handles, source/matrix.hpp:183-187)."""
    from gamba_b200.stepfile import to_indirect
    from oracle import oracle
    s = synthetic_step(600, 800, 0.05, p, seed=13, top_frac=0.85, share=6)
    want, _, _ = oracle.reduce(s)
    eng = engines(p)
    assert eng.reduce(s).same_as(want)
    assert eng.reduce(to_indirect(s, seed=2)).same_as(want)
    eng.stage(to_indirect(s, seed=3))
    assert eng.reduce_staged().same_as(want)


# ---- tensor-core product D' = D + M B (kernels_schur.cuh) ------------------------------------------

def schur_engine(p, **opts):
    from gamba_b200 import LinalgEngine
    eng = LinalgEngine(p)
    eng.set_option("schur_min_top", 1)     # default 1024: the small test steps would take the sparse path
    eng.set_option("block", 2)             # default 1: blocked elimination only for big steps with dense multipliers
    for k, v in opts.items():
        eng.set_option(k, v)
    return eng


@pytest.mark.parametrize("p", [251, 32003, 65521, 65537, 2147483647])
def test_schur_product_synthetic(built, p):
    """The B half of the elimination as a limb-split int8 product: 2 limbs for p < 2^16, 4 above; ragged
    tile edges (rows, columns and k not multiples of the tile), bit-exact against the oracle."""
    from oracle import oracle
    eng = schur_engine(p)
    for (nr, nc, dens, seed) in [(400, 520, 0.06, 5), (900, 1300, 0.03, 8), (333, 777, 0.05, 9)]:
        s = synthetic_step(nr, nc, dens, p, seed=seed, top_frac=0.85, share=7)
        want, fma_top, _ = oracle.reduce(s)
        assert eng.reduce(s).same_as(want), (p, nr, nc)
        c = eng.counters()
        assert c["schur_macs"] == s.n_bot * s.n_top * s.n_nonpiv and c["fma_top"] == fma_top
    eng.close()


@pytest.mark.parametrize("p", [32003, 2147483647])
def test_schur_accumulator_folds(built, p):
    """s32 accumulators are folded into D' every `kchunk` k-steps (the exactness bound is 256 / 128 k-steps of
    64): forced down to 1 and 3 here so that small steps take many folds."""
    from oracle import oracle
    s = synthetic_step(1200, 1500, 0.03, p, seed=17, top_frac=0.9, share=9)
    want, _, _ = oracle.reduce(s)
    for kc in (1, 3):
        eng = schur_engine(p, schur_kchunk=kc)
        assert eng.reduce(s).same_as(want), kc
        eng.close()


@pytest.mark.parametrize("block", [1, 0])
@pytest.mark.parametrize("name", ["cyclic7-15", "cyclic7-31", "kat9-16", "eco10-31", "cyclic7-8"])
def test_schur_every_step_matches_oracle(step_dirs, name, block):
    """block=1: R rows per CTA (kernels_elim_block.cuh); block=0: one row per CTA (kernels_elim.cuh) - both
    hand the multipliers to the tensor-core product."""
    d = step_dirs(name)
    s0 = load_step(sorted(glob.glob(os.path.join(d, "step_*.bin")))[0])
    eng = schur_engine(s0.p, block=block)
    assert check_dir(d, eng) > 0
    eng.close()


@pytest.mark.parametrize("name", ["cyclic8-15", "cyclic8-31", "kat10-15", "eco11-16"])
def test_blocked_elimination_larger_steps(step_dirs, name):
    """Several column tiles of 3392 columns and rows with different first pivots in one block of rows."""
    d = step_dirs(name, ("--dump-min-nnz", "100000"))
    s0 = load_step(sorted(glob.glob(os.path.join(d, "step_*.bin")))[0])
    eng = schur_engine(s0.p)
    assert check_dir(d, eng) >= 5
    assert eng.counters()["schur_macs"] > 0
    eng.close()


@pytest.mark.parametrize("tile", [64, 256])
def test_blocked_elimination_small_tiles(step_dirs, tile):
    eng = schur_engine(32003, tile_cols=tile)
    assert check_dir(step_dirs("cyclic7-15"), eng) > 0
    assert eng.counters()["n_tiles"] >= 4
    eng.close()


@pytest.mark.parametrize("name", ["cyclic8-15", "cyclic8-31", "kat10-15"])
def test_sparse_path_still_matches_oracle(step_dirs, name):
    """Option schur=0: both halves of the elimination in elim_kernel (the path for operands beyond schur_max_gb)."""
    from gamba_b200 import LinalgEngine
    d = step_dirs(name, ("--dump-min-nnz", "100000"))
    s0 = load_step(sorted(glob.glob(os.path.join(d, "step_*.bin")))[0])
    eng = LinalgEngine(s0.p)
    eng.set_option("schur", 0)
    assert check_dir(d, eng) >= 5
    assert eng.counters()["schur_macs"] == 0
    eng.close()


def test_schur_with_small_tiles(step_dirs):
    eng = schur_engine(32003, tile_cols=256)
    assert check_dir(step_dirs("cyclic7-15"), eng) > 0
    eng.close()


# ---- panel mode: blocked triangular solve, earlier pivots' contribution on the tensor cores ---------------

def panel_engine(p, panel_t=128, **opts):
    from gamba_b200 import LinalgEngine
    eng = LinalgEngine(p)
    eng.set_option("schur_min_top", 1)
    eng.set_option("panel", 2)             # default 1: only where the cost estimate from the previous step says it pays
    eng.set_option("panel_t", panel_t)     # default: 1024-3072, the test steps would be a single panel
    for k, v in opts.items():
        eng.set_option(k, v)
    return eng


@pytest.mark.parametrize("p", [251, 32003, 65521, 2147483647])
def test_panel_mode_synthetic(built, p):
    from oracle import oracle
    eng = panel_engine(p)
    for (nr, nc, dens, seed) in [(900, 1300, 0.03, 8), (1500, 1900, 0.02, 3), (333, 777, 0.05, 9)]:
        s = synthetic_step(nr, nc, dens, p, seed=seed, top_frac=0.85, share=7)
        want, fma_top, _ = oracle.reduce(s)
        assert eng.reduce(s).same_as(want), (p, nr, nc)
        c = eng.counters()
        assert c["fma_top"] == fma_top and c["n_tiles"] >= 3
    eng.close()


@pytest.mark.parametrize("name", ["cyclic7-15", "cyclic7-31", "kat9-16", "eco10-31", "cyclic8-15", "kat10-15"])
def test_panel_mode_every_step_matches_oracle(step_dirs, name):
    d = step_dirs(name) if name not in ("cyclic8-15", "kat10-15") else step_dirs(name, ("--dump-min-nnz", "100000"))
    s0 = load_step(sorted(glob.glob(os.path.join(d, "step_*.bin")))[0])
    eng = panel_engine(s0.p, panel_t=256 if name in ("cyclic8-15", "kat10-15") else 128)
    assert check_dir(d, eng) > 0
    eng.close()


def test_panel_mode_accumulator_folds(built):
    from oracle import oracle
    s = synthetic_step(1500, 1900, 0.02, 32003, seed=23, top_frac=0.9, share=9)
    want, _, _ = oracle.reduce(s)
    eng = panel_engine(32003, schur_kchunk=1)
    assert eng.reduce(s).same_as(want)
    eng.close()


# ---- Monte-Carlo combinations in front of the echelon ------------------------------------------------------

@pytest.mark.parametrize("p", [32003, 2147483647])
def test_mc_echelon_combinations_are_bit_exact(step_dirs, p):
    """K random combinations of the rows of D' (a tensor-core product) go into the echelon instead of the rows; the
    canonical RREF must be that of the rows. rank_hint far too small: the margin test must send the step back to the
    rows themselves; rank_hint = true rank: the combinations are accepted."""
    from gamba_b200 import LinalgEngine
    name = "cyclic8-15" if p == 32003 else "cyclic8-31"
    d = step_dirs(name, ("--dump-min-nnz", "100000"))
    eng = LinalgEngine(p)
    used = rejected = 0
    for f in sorted(glob.glob(os.path.join(d, "step_*.bin"))):
        if f.endswith("step_final.bin"):
            continue
        s = load_step(f)
        want = load_rows(f.replace("step_", "rows_"))
        if s.n_bot < 512:
            continue
        eng.set_option("rank_hint", want.n_new)
        assert eng.reduce(s).same_as(want), f
        used += eng.counters()["mc_ech_combinations"] > 0
        if want.n_new > 160:
            eng.set_option("rank_hint", 8)
            assert eng.reduce(s).same_as(want), f
            rejected += eng.counters()["mc_ech_combinations"] == 0
    assert used > 0 and rejected > 0
    eng.set_option("mc_ech", 0)
    eng.close()


# ---- block-sparse operands: the products skip the 128 x bn blocks of A / B that hold no entry ----------------------------

@pytest.mark.parametrize("p", [32003, 2147483647])
@pytest.mark.parametrize("mode", ["panel", "block"])
@pytest.mark.parametrize("dens", [0.004, 0.0002])
def test_block_sparse_operands_synthetic(built, p, mode, dens):
    """n_top >= 8192 switches the block-sparse operands on. Density 0.004: every block holds entries (full lists); 0.0002
    (2 entries per row): a quarter of the blocks is empty. Both with the blocked solve (A and B block-sparse) and with a sparse
    solve (B only), and the same with bsparse=0."""
    from gamba_b200 import LinalgEngine
    from oracle import oracle
    s = synthetic_step(10000, 11500, dens, p, seed=31, top_frac=0.88, share=20)
    want, fma_top, _ = oracle.reduce(s, threads=oracle.host_threads())
    for bsp in (1, 0):
        eng = LinalgEngine(p)
        eng.set_option("schur_min_top", 1); eng.set_option("bsparse", bsp)
        if mode == "panel":
            eng.set_option("panel", 2); eng.set_option("panel_t", 512)
        else:
            eng.set_option("panel", 0); eng.set_option("block", 2)
        assert eng.reduce(s).same_as(want), (p, mode, bsp, dens)
        c = eng.counters()
        assert c["fma_top"] == fma_top and c["solve_mode"] == (2 if mode == "panel" else 1)
        eng.close()
