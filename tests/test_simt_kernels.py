"""CPU-only checks of the KERNEL SOURCES: icp_rust_b200/csrc/*.cuh compiled with g++ against the
lock-step CTA emulator (tests/simt/) and compared with the oracle.  This validates barrier
placement, the exact selection and the ordered accumulation without a GPU; the -m gpu tests
repeat the comparison on the real sm_100a build through the C ABI."""
import numpy as np
import pytest

import oracle
from tests import refdata as R
from tests import simt_emul as E

I6 = np.array([1., 0., 0., 1., 0., 0.])


def _oracle_ref(srcs, dsts, T0s, max_iter):
    out = []
    for s, d, t in zip(srcs, dsts, T0s):
        out.append(oracle.icp_estimate(s, d, t, max_iter, nn_mode=oracle.NN_BRUTE, corr_mode=oracle.CORR_ALL))
    return out


def _compare(g, refs, srcs, max_iter, ordered=True):
    """ordered sums: everything bit-identical.  Tree reductions: indices and iteration counts
    identical, transforms within the 1e-9 contract of BASELINE.json."""
    so = g["src_off"]
    for b, o in enumerate(refs):
        n = len(srcs[b])
        gc = g["corr"][int(so[b]) * max_iter:int(so[b]) * max_iter + n * max_iter].reshape(max_iter, n)
        assert np.array_equal(gc, o["corr"]), f"pair {b}: correspondences differ"
        assert np.array_equal(g["inner"][b], o["inner"]), f"pair {b}: inner iteration counts differ"
        if ordered is True:
            assert np.array_equal(g["T"][b], o["T"]), f"pair {b}: transform differs"
        else:
            assert np.abs(g["T"][b] - o["T"]).max() <= 1e-9, f"pair {b}: transform differs by more than 1e-9"
        assert int(g["corr_hash"][b]) == int(o["corr_hash"]), f"pair {b}: hash differs"


@pytest.mark.parametrize("threads", [32, 128, 256])
@pytest.mark.parametrize("mode", [0, 1])
def test_calc_stddevs_exact(threads, mode):
    rng = np.random.default_rng(5)
    sets = [rng.normal(0, 3, (n, 2)) for n in (1, 2, 3, 10, 11, 64, 665, 666, 1080)]
    sets.append(np.zeros((50, 2)))                              # MAD == 0
    sets.append(np.repeat(rng.normal(0, 1, (5, 2)), 40, axis=0))  # heavy duplicates
    sets.append(rng.normal(0, 1e-300, (33, 2)))                 # denormal-scale
    sets.append(rng.normal(0, 1e150, (34, 2)))
    sets.append(np.array(R.MEASUREMENTS_30))
    sig, med = E.calc_stddevs(sets, mode=mode, threads=threads)
    for k, s in enumerate(sets):
        exp = oracle.calc_stddevs(s)
        assert np.array_equal(sig[k], exp), (k, sig[k], exp)
        for a in range(2):
            assert med[k, a] == oracle.median(s[:, a])


def test_calc_stddevs_reference_vectors():
    # stats.rs:69-102 medians / MADs through the device selection
    for vals, exp in R.MEDIAN_CASES:
        if exp is None:
            continue
        s = np.stack([vals, vals], axis=1)
        _, med = E.calc_stddevs([s])
        assert med[0, 0] == exp and med[0, 1] == exp
    for vals, exp in R.MAD_CASES:
        s = np.stack([vals, vals], axis=1)
        sig, _ = E.calc_stddevs([s])
        assert sig[0, 0] == 1.482602218505602 * exp


def _l_shape_case(dim):
    true_T = oracle.transform_new(R.L_TRUE_PARAM)
    init = oracle.transform_new(np.array(R.L_TRUE_PARAM) + np.array(R.L_NOISE_PARAM))
    if dim == 2:
        src = R.L_SHAPE_2D
        dst = np.array([oracle.transform_apply(true_T, p) for p in src])
    else:
        src = R.L_SHAPE_3D
        dst = np.array([oracle.transform_xy(true_T, p) for p in src])
    return src, dst, init, true_T


@pytest.mark.parametrize("ordered", [True, False, "lean"])
@pytest.mark.parametrize("dim", [2, 3])
@pytest.mark.parametrize("variant", [2, 3])
def test_fused_l_shape(dim, variant, ordered):  # lib.rs:510-595
    if ordered == "lean" and variant == 3:
        pytest.skip("icp_lean.cuh needs two service warps")
    src, dst, init, true_T = _l_shape_case(dim)
    g = E.icp_fused([src], [dst], [init], max_iter=20, variant=variant, ordered=ordered)
    _compare(g, _oracle_ref([src], [dst], [init], 20), [src], 20, ordered)
    for p, q in zip(src, dst):
        got = oracle.transform_apply(g["T"][0], p[:2])
        assert np.linalg.norm(got - q[:2]) < 1e-3


@pytest.mark.parametrize("ordered", [True, False, "lean"])
@pytest.mark.parametrize("schedule", [0, 1, 7])
def test_fused_real_scans(scans2d, schedule, ordered):
    E.set_schedule(schedule)
    try:
        srcs = [scans2d[1], scans2d[98]]
        dsts = [scans2d[2], scans2d[99]]   # 099 holds duplicated (0,0) points: index ties
        T0 = [I6, I6]
        g = E.icp_fused(srcs, dsts, T0, max_iter=3, variant=1, grid=2, ordered=ordered)
        _compare(g, _oracle_ref(srcs, dsts, T0, 3), srcs, 3, ordered)
        if ordered is not True:   # tree reductions are schedule-independent: same bits under every interleaving
            # (the lean schedule sums factorised terms, icp_lean.cuh: its bits differ from icp_fused.cuh's
            # tree mode, by rounding only)
            E.set_schedule(0)
            ref = E.icp_fused(srcs, dsts, T0, max_iter=3, variant=1, grid=1, ordered=ordered)
            assert np.array_equal(ref["T"], g["T"])
            other = E.icp_fused(srcs, dsts, T0, max_iter=3, variant=1, grid=1, ordered=False)
            assert np.abs(other["T"] - g["T"]).max() <= 1e-11
    finally:
        E.set_schedule(0)


def test_fused_degenerate_pairs():
    rng = np.random.default_rng(3)
    a = rng.normal(0, 100, (40, 2))
    srcs = [a, a[:1], np.zeros((0, 2)), a, np.zeros((30, 2))]
    dsts = [a.copy(), a, a, a[:1], np.zeros((30, 2))]  # identical clouds: sigma == 0 -> None
    T0 = [I6] * 5
    for ordered in (True, False, "lean"):
        g = E.icp_fused(srcs, dsts, T0, max_iter=4, variant=2, grid=1, ordered=ordered)
        _compare(g, _oracle_ref(srcs, dsts, T0, 4), srcs, 4, ordered)
        assert np.array_equal(g["T"][0], I6)
        assert (g["inner"][0] >> 16 == oracle.BREAK_NONE).all()


def test_fused_status_codes():
    rng = np.random.default_rng(4)
    a = rng.normal(0, 100, (20, 2))
    bad = a.copy()
    bad[3, 1] = np.nan
    srcs = [a, bad, a]
    dsts = [np.zeros((0, 2)), a, a]
    T0 = [I6, I6, np.array([1., 0., 0., 1., np.inf, 0.])]
    for ordered in (True, "lean"):
        g = E.icp_fused(srcs, dsts, T0, max_iter=2, variant=2, ordered=ordered)
        assert list(g["status"]) == [1, 3, 3]
        assert g["err"] == 3
        for b in range(3):
            assert np.array_equal(g["T"][b], T0[b], equal_nan=True)


def test_lean_crowded_and_odd_sizes():
    """icp_lean.cuh: the short selection path must hand over to the general selection whenever it
    cannot decide (crowded bins, heavy duplicates, tiny spreads around a large offset) and the
    result must not depend on which path ran.  Odd and even n (stats.rs:23-27)."""
    rng = np.random.default_rng(12)
    srcs, dsts = [], []
    for n, spread in ((1101, 0.5), (1100, 1e-7), (333, 30.0), (64, 3.0)):
        d = rng.uniform(-1000, 1000, (n, 2))
        s = d + rng.normal(0, spread, (n, 2)) + np.array([0.3, -0.2])
        s[: n // 3] = d[: n // 3] + np.array([0.25, 0.25])   # one third share the same residual
        srcs.append(s); dsts.append(d)
    T0 = [I6] * len(srcs)
    g = E.icp_fused(srcs, dsts, T0, max_iter=3, variant=1, grid=2, ordered="lean")
    _compare(g, _oracle_ref(srcs, dsts, T0, 3), srcs, 3, ordered=False)
    ref = E.icp_fused(srcs, dsts, T0, max_iter=3, variant=1, grid=1, ordered=False)
    assert np.abs(ref["T"] - g["T"]).max() <= 1e-11 and np.array_equal(ref["inner"], g["inner"])


# ------------------------------------------------------------------ streaming (large-scan) kernel
@pytest.mark.parametrize("dim", [2, 3])
@pytest.mark.parametrize("threads", [32, 128])
def test_stream_l_shape(dim, threads):  # lib.rs:510-595 through icp_stream.cuh
    src, dst, init, _ = _l_shape_case(dim)
    g = E.icp_stream([src], [dst], [init], max_iter=20, threads=threads)
    _compare(g, _oracle_ref([src], [dst], [init], 20), [src], 20, ordered=False)


@pytest.mark.parametrize("nn_group", [False, True])
@pytest.mark.parametrize("schedule", [0, 7])
def test_stream_real_scans_and_degenerates(scans2d, schedule, nn_group):
    E.set_schedule(schedule)
    E.set_stream_nn_group(nn_group)
    try:
        rng = np.random.default_rng(3)
        a = rng.normal(0, 100, (40, 2))
        srcs = [scans2d[1], scans2d[98], a, a[:1], np.zeros((0, 2)), np.zeros((30, 2))]
        dsts = [scans2d[2], scans2d[99], a.copy(), a, a, np.zeros((30, 2))]
        T0 = [I6] * len(srcs)
        g = E.icp_stream(srcs, dsts, T0, max_iter=3, threads=64, grid=2)
        _compare(g, _oracle_ref(srcs, dsts, T0, 3), srcs, 3, ordered=False)
        # same bits as the fused kernel's warp-tree mode?  Not required (different summation
        # tree); but statuses and counts are.
        assert (g["status"] == 0).all()
        if nn_group:   # 3-D clouds through the group search as well
            srcs3, dsts3 = _noisy_pairs(2, 400, 17, dim=3)
            g = E.icp_stream(srcs3, dsts3, [I6] * 2, max_iter=3, threads=64, grid=1)
            _compare(g, _oracle_ref(srcs3, dsts3, [I6] * 2, 3), srcs3, 3, ordered=False)
    finally:
        E.set_schedule(0)
        E.set_stream_nn_group(False)


@pytest.mark.parametrize("dim", [2, 3])
def test_stream_far_points_take_the_warp_search(dim):
    """Source points without a counterpart (their match is far: the seeded ball covers many rows) are
    searched by a whole warp after the per-thread pass; same indices, counts and transforms."""
    srcs, dsts = _noisy_pairs(2, 500, 41, dim=dim)
    rng = np.random.default_rng(9)
    for s_ in srcs:
        far = rng.choice(len(s_), 12, replace=False)
        s_[far, :2] += rng.uniform(120, 260, (12, 2)) * rng.choice([-1, 1], (12, 2))
    T0 = [I6] * 2
    g = E.icp_stream(srcs, dsts, T0, max_iter=4, threads=64, grid=2)
    _compare(g, _oracle_ref(srcs, dsts, T0, 4), srcs, 4, ordered=False)


def test_stream_crowded_bins_zoom():
    """3000 residuals with a tiny spread around a large offset: the FAST map's bins are crowded
    (zoom) or unresolvable in binary32 (RADIX fallback) - the medians must still be exact."""
    rng = np.random.default_rng(11)
    n = 3000
    d = rng.uniform(-1000, 1000, (n, 2))
    s = d + rng.normal(0, 0.5, (n, 2)) + np.array([0.3, -0.2])
    g = E.icp_stream([s], [d], [I6], max_iter=2, threads=128)
    _compare(g, _oracle_ref([s], [d], [I6], 2), [s], 2, ordered=False)


def test_stream_status_codes():
    rng = np.random.default_rng(4)
    a = rng.normal(0, 100, (20, 2))
    bad = a.copy()
    bad[3, 1] = np.nan
    g = E.icp_stream([a, bad], [np.zeros((0, 2)), a], [I6, I6], max_iter=2, threads=32)
    assert list(g["status"]) == [1, 3] and g["err"] == 3


# ------------------------------------------------------------------------------------------------
# Shared-target search (icp_target.cuh + icp_target_nn.cuh): KdTree::search, lib.rs:121,164
def _map_queries(seed, count=6, beams=360):
    import icp_rust_b200 as icp
    scans, off, T0 = icp.synth_map_scans(count, beams=beams, seed=seed, nthreads=1)
    out = []
    for b in range(count):
        s = scans[off[b]:off[b + 1]]
        R2 = T0[b, :4].reshape(2, 2)
        out.append(s @ R2.T + T0[b, 4:6])
    return np.concatenate(out)


@pytest.mark.parametrize("cells_per_point,fast", [(2, True), (32, True), (200, True), (32, "warp"), (32, "group4"), (8, "group8")])
def test_target_search_fast_path_is_exact(cells_per_point, fast):
    """Seeded, binary32-filtered disc search == brute force on a map sampled densely along walls:
    near-ties (neighbours on a wall), exact duplicates, queries on target points, far outside."""
    import icp_rust_b200 as icp
    m = icp.synth_map(6000, seed=7)
    m[700:720] = m[100]              # duplicates: lowest index must win
    m[3000] = m[2999]
    rng = np.random.default_rng(5)
    q = np.concatenate([_map_queries(7), m[95:105], m[2995:3005], m[::97] + 1e-9,
                        rng.uniform(-5e3, 45e3, (300, 2)), rng.uniform(-1e7, 1e7, (16, 2)),
                        np.zeros((1, 2)), [[1e300, -1e300], [5e-324, 0.0]]])
    exp = oracle.nn_search(m, q, oracle.NN_BRUTE)
    got = E.target_search(m, q, budget=len(m) * cells_per_point, fast=fast)
    assert np.array_equal(got, exp)
    assert np.array_equal(E.target_search(m, q[:400], budget=len(m) * 2, fast=False), exp[:400])
    if fast != True:      # a query count that is not a multiple of 32, and fewer than 32
        for cut in (len(q) - 7, 5, 1):
            assert np.array_equal(E.target_search(m, q[:cut], budget=len(m) * cells_per_point, fast=fast), exp[:cut])


@pytest.mark.parametrize("fast", [True, "warp", "group4"])
def test_target_search_fast_path_odd_geometries(fast):
    rng = np.random.default_rng(9)
    line = np.stack([np.linspace(-1.0, 1.0, 500), np.full(500, 3.25)], axis=1)       # flat in y
    lattice = np.stack(np.meshgrid(np.arange(40.), np.arange(40.)), -1).reshape(-1, 2)  # ties everywhere
    tiny = rng.normal(0, 1e-7, (300, 2)) + 1e3                                        # far from the origin
    one = np.array([[2.0, -1.0]])
    for d in (line, lattice, tiny, one):
        span = np.ptp(d, axis=0).max() + 1.0
        q = np.concatenate([d[:: max(1, len(d) // 50)], d.mean(0) + rng.normal(0, span, (200, 2)),
                            d.mean(0) + rng.normal(0, span * 1e-3, (100, 2))])
        if d is lattice:
            q = np.concatenate([q, lattice[:200] + 0.5, lattice[:100] + [0.5, 0.0]])   # 4-way and 2-way ties
        exp = oracle.nn_search(d, q, oracle.NN_BRUTE)
        for cpp in (4, 64):
            assert np.array_equal(E.target_search(d, q, budget=max(64, len(d) * cpp), fast=fast), exp)


@pytest.mark.parametrize("dim", [2, 3])
def test_seeded_disc_walk_is_exact_for_any_seed(dim):
    """grid_nn_search_seeded / _auto: whatever point seeds the disc (the true neighbour, a far point,
    none), whatever the cell size, scale and offset of the cloud, the result is the brute-force one."""
    rng = np.random.default_rng(100 + dim)
    for trial in range(14):
        m = int(rng.integers(1, 1500))
        scale = 10.0 ** rng.uniform(-6, 6)
        offset = rng.normal(0, 1, dim) * scale * 10.0 ** rng.uniform(-1, 3)
        aniso = 10.0 ** rng.uniform(-2, 0, dim)
        kind = trial % 4
        if kind == 0:
            d = rng.normal(0, 1, (m, dim))
        elif kind == 1:                                   # points on a curve (scan-like), dense
            t = np.sort(rng.uniform(0, 6.28, m))
            d = np.stack([np.cos(t), np.sin(2 * t)] + ([0.1 * np.sin(5 * t)] if dim == 3 else []), axis=1)
            d += rng.normal(0, 1e-3, d.shape)
        elif kind == 2:                                   # lattice: exact ties
            side = int(np.ceil(m ** (1.0 / dim)))
            d = np.stack(np.meshgrid(*[np.arange(float(side))] * dim), -1).reshape(-1, dim)[:m]
        else:                                             # clusters with duplicates
            d = rng.normal(0, 1, (m, dim))
            d[rng.integers(0, m, m // 3)] = d[rng.integers(0, m, 1)]
        d = d * aniso * scale + offset
        m = len(d)
        span = np.ptp(d, axis=0).max() + scale
        q = np.concatenate([d[rng.integers(0, m, 120)] + rng.normal(0, span * 1e-3, (120, dim)),
                            d[rng.integers(0, m, 40)],
                            d.mean(0) + rng.normal(0, span, (60, dim)),
                            d[rng.integers(0, m, 30)] + ([0.5] * dim if kind == 2 else [0.0] * dim) * np.array(aniso) * scale])
        exp = oracle.nn_search(d, q, oracle.NN_BRUTE)
        seeds = np.concatenate([exp[:60], rng.integers(0, m, 130), np.full(len(q) - 190, -1)])
        budget = int(max(8, m * 10.0 ** rng.uniform(-1, 1.5)))
        got = E.grid_seeded_search(d, q, seeds, min(budget, 1 << 20))
        assert np.array_equal(got, exp), f"trial {trial} dim {dim} m {m} budget {budget}"


@pytest.mark.parametrize("fast", [True, "warp", "group4"])
def test_target_search_random_geometries(fast):
    """Seeded disc + binary32 filter on shared 2-D targets across scales, offsets and cell sizes."""
    rng = np.random.default_rng(77)
    for trial in range(10):
        m = int(rng.integers(2, 3000))
        scale = 10.0 ** rng.uniform(-5, 7)
        offset = rng.normal(0, 1, 2) * scale * 10.0 ** rng.uniform(-1, 4)
        t = np.sort(rng.uniform(0, 6.28, m))
        d = np.stack([np.cos(t) * (1 + 0.3 * np.sin(7 * t)), np.sin(t)], axis=1) + rng.normal(0, 10.0 ** rng.uniform(-6, -2), (m, 2))
        d[rng.integers(0, m, m // 10)] = d[rng.integers(0, m, 1)]
        d = d * scale + offset
        q = np.concatenate([d[rng.integers(0, m, 150)] * (1 + rng.normal(0, 1e-3, (150, 1))), d[rng.integers(0, m, 50)],
                            d.mean(0) + rng.normal(0, 3 * scale, (80, 2))])
        exp = oracle.nn_search(d, q, oracle.NN_BRUTE)
        cpp = float(10.0 ** rng.uniform(0, 2))
        assert np.array_equal(E.target_search(d, q, budget=max(64, int(m * cpp)), fast=fast), exp), f"trial {trial}"


# ------------------------------------------------------------------------------------------------
# Knife-edge guard (GuardBands, icp_fused.cuh) + ordered re-run, final cost, ordered streaming sums
def _noisy_pairs(B, n, seed, dim=2):
    rng = np.random.default_rng(seed)
    srcs, dsts = [], []
    for _ in range(B):
        a = rng.uniform(-300, 300, (n, dim))
        if dim == 3:
            a[:, 2] *= 0.1
        T = oracle.transform_new([rng.uniform(-4, 4), rng.uniform(-4, 4), rng.uniform(-0.01, 0.01)])
        d = a.copy()
        d[:, :2] = a[:, :2] @ T[:4].reshape(2, 2).T + T[4:]
        d += rng.normal(0, 0.7, d.shape)
        srcs.append(a)
        dsts.append(d[rng.permutation(n)])
    return srcs, dsts


@pytest.mark.parametrize("mode", [False, "lean"])
def test_knife_edge_guard_forced_band(mode):
    """Guard bands blown up by 1e9 (every decision is 'inside the band'): every pair that takes a
    decision is flagged and re-run with ordered sums, so the whole batch becomes BIT-identical to the
    oracle although the first pass used tree reductions.  With the real bands nothing on this
    batch is flagged and the tree results stand."""
    srcs, dsts = _noisy_pairs(5, 150, 21)
    T0 = [I6] * 5
    refs = _oracle_ref(srcs, dsts, T0, 4)
    try:
        E.set_options(guard_scale=1e9)
        g = E.icp_fused(srcs, dsts, T0, 4, variant=0, grid=2, ordered=mode)
        assert E.last_redo_count() == 5
        _compare(g, refs, srcs, 4, ordered=True)
        E.set_options(guard_scale=1.0)
        g = E.icp_fused(srcs, dsts, T0, 4, variant=0, grid=2, ordered=mode)
        assert E.last_redo_count() == 0
        _compare(g, refs, srcs, 4, ordered=False)
    finally:
        E.set_options()


def test_knife_edge_guard_partial_band():
    """A band between the pairs' own margins: only the pairs whose tightest decision lies inside it
    are re-run (bit-identical); the others keep their tree result (within 1e-9)."""
    srcs, dsts = _noisy_pairs(8, 120, 33)
    T0 = [I6] * 8
    refs = _oracle_ref(srcs, dsts, T0, 5)
    so = np.zeros(9, np.uint32); so[1:] = np.cumsum([len(s) for s in srcs])
    o = oracle.icp_estimate_batch(2, np.concatenate(srcs), so, np.concatenate(dsts), so, np.array(T0), 5,
                                  nn_mode=oracle.NN_BRUTE, want_margins=True)
    err_m = np.sort(o["margins"][:, 1])
    band = float(np.sqrt(err_m[3] * err_m[4]))          # half of the pairs have an error test inside it
    try:
        E.set_options(guard_scale=band / 1e-11, delta_band_scale=1e-11 / band)   # the delta band keeps its own width
        g = E.icp_fused(srcs, dsts, T0, 5, variant=0, grid=3, ordered="lean")
        flagged = E.last_redo_count()
    finally:
        E.set_options()
    assert 4 <= flagged < 8     # the det band (scaled alike) may add pairs, never remove any
    _compare(g, refs, srcs, 5, ordered=False)
    exact = sum(np.array_equal(g["T"][b], refs[b]["T"]) for b in range(8))
    assert exact >= flagged


@pytest.mark.parametrize("dim", [2, 3])
def test_streaming_kernel_ordered_sums(dim):
    """Reference-order sums in the large-scan kernel (tiles of 256 points through shared memory, the
    chain's accumulators carried across tiles): bit-identical to the oracle, 1 to 3 tiles + ragged."""
    srcs, dsts = _noisy_pairs(3, 300, 5, dim=dim)
    srcs[1], dsts[1] = srcs[1][:256], dsts[1][:256]      # exactly one tile
    srcs[2], dsts[2] = np.concatenate([srcs[2], srcs[0] + 3.0]), np.concatenate([dsts[2], dsts[0] + 3.0])  # 600 -> 3 tiles
    T0 = [I6] * 3
    refs = _oracle_ref(srcs, dsts, T0, 3)
    try:
        E.set_options(stream_ordered=True, want_cost=3)
        for threads in (32, 64):
            g = E.icp_stream(srcs, dsts, T0, 3, threads=threads, grid=2)
            _compare(g, refs, srcs, 3, ordered=True)
            assert np.array_equal(E.last_cost(), [r["cost"] for r in refs])
        # and as the re-run path of the guard
        E.set_options(guard_scale=1e9)
        g = E.icp_stream(srcs, dsts, T0, 3, threads=64, grid=1)
        assert E.last_redo_count() == 3
        _compare(g, refs, srcs, 3, ordered=True)
    finally:
        E.set_options()


@pytest.mark.parametrize("mode", [True, False, "lean"])
def test_final_cost_output(mode):
    """out_cost = huber_error (lib.rs:45-50) of the last inner solve: bit-identical with ordered sums,
    ~1e-13 relative with the tree; 0 for max_iter == 0; degenerate pairs (n < 2)."""
    srcs, dsts = _noisy_pairs(4, 90, 8)
    srcs.append(srcs[0][:1]); dsts.append(dsts[0])
    T0 = [I6] * 5
    refs = _oracle_ref(srcs, dsts, T0, 3)
    try:
        E.set_options(want_cost=5)
        E.icp_fused(srcs, dsts, T0, 3, variant=0, grid=2, ordered=mode)
        cost = E.last_cost().copy()
        E.set_options(want_cost=5)
        E.icp_fused(srcs, dsts, T0, 0, variant=0, grid=2, ordered=mode)
        assert np.array_equal(E.last_cost(), np.zeros(5))
    finally:
        E.set_options()
    exp = np.array([r["cost"] for r in refs])
    if mode is True:
        assert np.array_equal(cost, exp)
    else:
        assert np.allclose(cost, exp, rtol=1e-12, atol=0)


def test_shared_source_batch():
    """ICP_B200_FLAG_SHARED_SOURCE: one scan, several initial transforms — same results as the scan
    repeated per pair."""
    srcs, dsts = _noisy_pairs(1, 140, 12)
    inits = [I6, oracle.transform_new([2., -1., 0.004]), oracle.transform_new([-3., 2., -0.006])]
    refs = _oracle_ref(srcs * 3, dsts * 3, inits, 4)
    g_rep = E.icp_fused(srcs * 3, dsts * 3, inits, 4, variant=0, grid=2, ordered=True)
    _compare(g_rep, refs, srcs * 3, 4, ordered=True)
    try:
        E.set_options(shared_src=True)
        g = E.icp_fused(srcs * 3, dsts * 3, inits, 4, variant=0, grid=2, ordered=True)   # src_off[0:2] is the one scan
        for k in ("T", "inner", "corr", "corr_hash"):
            assert np.array_equal(g[k], g_rep[k]), k
        g = E.icp_fused(srcs * 3, dsts * 3, inits, 4, variant=0, grid=2, ordered="lean")
        _compare(g, refs, srcs * 3, 4, ordered=False)
    finally:
        E.set_options()


def test_lean_dedicated_service_warps(scans2d):
    """icp_lean.cuh with SVC: warps 0 and 1 own no points (they serve the medians while the workers
    accumulate).  Same checks as the other schedules: real scans with index ties, the L shape,
    degenerate pairs, crowded / odd sizes, the guard's re-run and the final cost; every schedule."""
    srcs = [scans2d[1], scans2d[98]]
    dsts = [scans2d[2], scans2d[99]]
    T0 = [I6, I6]
    refs = _oracle_ref(srcs, dsts, T0, 3)
    outs = []
    for schedule in (0, 1, 7):
        E.set_schedule(schedule)
        try:
            g = E.icp_fused(srcs, dsts, T0, max_iter=3, variant=4, grid=2, ordered="svc")
        finally:
            E.set_schedule(0)
        _compare(g, refs, srcs, 3, ordered=False)
        outs.append(g["T"])
    assert np.array_equal(outs[0], outs[1]) and np.array_equal(outs[0], outs[2])
    for dim in (2, 3):
        src, dst, init, _ = _l_shape_case(dim)
        g = E.icp_fused([src], [dst], [init], max_iter=20, variant=0, ordered="svc")
        _compare(g, _oracle_ref([src], [dst], [init], 20), [src], 20, ordered=False)
    rng = np.random.default_rng(3)
    a = rng.normal(0, 100, (40, 2))
    ds, dd = [a, a[:1], np.zeros((0, 2)), a, np.zeros((30, 2))], [a.copy(), a, a, a[:1], np.zeros((30, 2))]
    g = E.icp_fused(ds, dd, [I6] * 5, max_iter=4, variant=0, grid=1, ordered="svc")
    _compare(g, _oracle_ref(ds, dd, [I6] * 5, 4), ds, 4, ordered=False)
    assert np.array_equal(g["T"][0], I6)
    rng = np.random.default_rng(12)
    cs, cd = [], []
    for n, spread in ((1101, 0.5), (1100, 1e-7), (333, 30.0), (64, 3.0)):
        d = rng.uniform(-1000, 1000, (n, 2))
        s_ = d + rng.normal(0, spread, (n, 2)) + np.array([0.3, -0.2])
        s_[: n // 3] = d[: n // 3] + np.array([0.25, 0.25])
        cs.append(s_); cd.append(d)
    g = E.icp_fused(cs, cd, [I6] * 4, max_iter=3, variant=4, grid=2, ordered="svc")
    crefs = _oracle_ref(cs, cd, [I6] * 4, 3)
    _compare(g, crefs, cs, 3, ordered=False)
    try:
        E.set_options(guard_scale=1e9, want_cost=4)
        g = E.icp_fused(cs, cd, [I6] * 4, max_iter=3, variant=4, grid=2, ordered="svc")
        assert E.last_redo_count() == 4
        _compare(g, crefs, cs, 3, ordered=True)
        assert np.array_equal(E.last_cost(), [r["cost"] for r in crefs])
    finally:
        E.set_options()


@pytest.mark.parametrize("dim", [2, 3])
def test_outer_fixed_point_exit(dim):
    """icp_fused.cuh fused_repeats: an inner solve that accepts no update leaves the outer transform
    unchanged, so every later outer iteration repeats it exactly; the kernels write those
    iterations' outputs instead of recomputing them.  Every output (transform, inner-iteration
    words, correspondences of ALL iterations or of the last one, checksum, final cost) equals the
    run that computes every iteration (ICP_B200_FIXED_POINT=0) bit for bit, and the oracle's."""
    srcs, dsts = _noisy_pairs(3, 80, 21, dim=dim)
    srcs.append(srcs[0][:1]); dsts.append(dsts[0])       # n < 2: the first iteration already repeats
    T0 = [I6] * 4
    max_iter = 12
    refs = _oracle_ref(srcs, dsts, T0, max_iter)
    first_idle = [int(np.argmax((r["inner"] & 0xffff) == 0)) for r in refs]
    assert all(((r["inner"] & 0xffff) == 0).any() for r in refs) and max(first_idle) < max_iter - 2, \
        "the test data must reach the fixed point early"
    runs = [("fused", dict(variant=0, grid=2, ordered=True)), ("fused", dict(variant=0, grid=2, ordered=False)),
            ("fused", dict(variant=0, grid=2, ordered="lean")), ("fused", dict(variant=0, grid=2, ordered="svc")),
            ("stream", dict(threads=64, grid=2))]
    for kind, kw in runs:
        fn = E.icp_fused if kind == "fused" else E.icp_stream
        out = {}
        for on in (True, False):
            try:
                E.set_fixed_point(on)
                E.set_options(want_cost=4)
                g = fn(srcs, dsts, T0, max_iter, **kw)
                g["cost"] = E.last_cost().copy()
                E.set_options(want_cost=4)
                g["corr_last"] = fn(srcs, dsts, T0, max_iter, corr_mode=1, **kw)["corr"]
            finally:
                E.set_fixed_point(True)
                E.set_options()
            out[on] = g
        for k in ("T", "inner", "corr", "corr_hash", "cost", "status", "corr_last"):
            assert np.array_equal(out[True][k], out[False][k]), (kind, kw, k)
        _compare(out[True], refs, srcs, max_iter, ordered=(kw.get("ordered") is True))
        so = out[True]["src_off"]
        for b, r in enumerate(refs):   # corr_mode LAST holds the last iteration's correspondences
            assert np.array_equal(out[True]["corr_last"][int(so[b]):int(so[b + 1])], r["corr"][-1])


def test_stream_3d_binary32_filter_ties_and_duplicates():
    """grid_nn_search_seeded_f32 (icp_target_nn.cuh): the streaming kernel's 3-D searches rank the
    candidates on binary32 records first; exact ties (two target points mirrored about a source
    point), duplicated target points (same coordinates, different index), near-ties below the
    filter's resolution and a cloud far from the origin must come out as the oracle's brute force
    has them (lowest index on ties), with the filter on and off."""
    rng = np.random.default_rng(77)
    for shift in (0.0, 4096.0):
        dst = rng.uniform(0, 10, (500, 3))
        dst[:, 2] *= 0.3
        src = rng.uniform(0.5, 9.5, (160, 3))
        src[:, 2] *= 0.3
        extra = []
        for k in range(0, 40):            # exact ties: mirrored pairs at a power-of-two offset
            d = 2.0 ** -(3 + k % 4)
            axis = k % 3
            e = np.zeros(3); e[axis] = d
            extra += [src[k] + e, src[k] - e]
        for k in range(40, 60):           # near-ties: 1e-9 apart in distance
            e = np.array([0.05, 0.0, 0.0])
            extra += [src[k] + e, src[k] - e * (1 + 1e-9)]
        dup = dst[:30].copy()             # duplicates of existing points (higher indices lose)
        dst = np.concatenate([dst, np.array(extra), dup]) + shift
        src = src + shift
        perm = rng.permutation(len(dst))
        dst = dst[perm]
        refs = _oracle_ref([src], [dst], [I6], 3)
        outs = []
        for on in (True, False):
            try:
                E.set_stream_nn_f32(on)
                g = E.icp_stream([src], [dst], [I6], 3, threads=64, grid=1)
            finally:
                E.set_stream_nn_f32(True)
            _compare(g, refs, [src], 3, ordered=False)
            outs.append(g)
        for k in ("T", "inner", "corr", "corr_hash"):
            assert np.array_equal(outs[0][k], outs[1][k]), k
