"""Parity of the sm_100a CUDA path (through the C ABI of include/icp_b200.h) with the CPU oracle.
Bar (BASELINE.json north_star): correspondence indices and inner iteration counts / break reasons
BIT-EXACT in both reduction modes; transforms bit-exact with ICP_B200_FLAG_ORDERED_SUMS (the
kernels then reproduce the reference's operation order) and within 1e-9 (absolute, per component:
radians-scale rotation entries and input length units) with the default warp-tree reductions."""
import os

import numpy as np
import pytest

import oracle
from tests import refdata as R

pytestmark = pytest.mark.gpu

I6 = np.array([1., 0., 0., 1., 0., 0.])


@pytest.fixture(scope="module")
def icp():
    import icp_rust_b200 as m
    return m


@pytest.fixture(scope="module")
def ctx(icp):
    c = icp.Context(0)
    yield c
    c.close()


def _cat(arrs, dim):
    off = np.zeros(len(arrs) + 1, dtype=np.uint32)
    off[1:] = np.cumsum([len(a) for a in arrs])
    flat = np.concatenate([np.asarray(a, dtype=np.float64).reshape(-1, dim) for a in arrs] + [np.zeros((0, dim))])
    return flat, off


def _run(ctx, icp, srcs, dsts, T0s, max_iter, dim=2, **kw):
    src, so = _cat(srcs, dim)
    dst, do = _cat(dsts, dim)
    g = ctx.estimate_batch(src, so, dst, do, np.array(T0s).reshape(-1, 6), max_iter, dim=dim,
                           corr_mode=icp.CORR_ALL, **kw)
    g["src_off"] = so
    return g


TOL = 1e-9  # north_star: "final transforms must agree within 1e-9 rad and 1e-9 m"


def _check_vs_oracle(g, srcs, dsts, T0s, max_iter, nn_mode=oracle.NN_KDTREE, ordered=True):
    so = g["src_off"]
    for b in range(len(srcs)):
        o = oracle.icp_estimate(srcs[b], dsts[b], T0s[b], max_iter, nn_mode=nn_mode, corr_mode=oracle.CORR_ALL)
        n = len(srcs[b])
        gc = g["corr"][int(so[b]) * max_iter:int(so[b]) * max_iter + n * max_iter].reshape(max_iter, n)
        assert np.array_equal(gc, o["corr"]), f"pair {b}: correspondence indices differ"
        assert np.array_equal(g["inner"][b], o["inner"]), f"pair {b}: inner iteration counts differ"
        if ordered:
            assert np.array_equal(g["T"][b], o["T"]), f"pair {b}: transform differs by {np.abs(g['T'][b] - o['T']).max()}"
        else:
            assert np.abs(g["T"][b] - o["T"]).max() <= TOL, f"pair {b}: transform differs by {np.abs(g['T'][b] - o['T']).max()}"
        assert int(g["corr_hash"][b]) == int(o["corr_hash"])
        assert g["status"][b] == 0


# ------------------------------------------------------------------------------------------------
def test_native_library_is_the_cuda_build(icp, ctx):
    assert os.path.exists(icp._lib.CUDA_SO)
    assert icp.cuda_lib().icp_b200_abi_version() == icp.ABI_VERSION


def test_calc_stddevs_exact(ctx):  # stats.rs:49-60 on the device selection
    rng = np.random.default_rng(5)
    sets = [rng.normal(0, 3, (n, 2)) for n in (1, 2, 3, 10, 11, 64, 665, 666, 1080, 1081, 3000)]
    sets.append(np.zeros((50, 2)))
    sets.append(np.repeat(rng.normal(0, 1, (5, 2)), 40, axis=0))
    sets.append(rng.normal(0, 1e-300, (33, 2)))
    sets.append(rng.normal(0, 1e150, (34, 2)))
    sets.append(np.array(R.MEASUREMENTS_30))
    for vals, exp in R.MEDIAN_CASES + R.MAD_CASES:
        if exp is not None:
            sets.append(np.stack([vals, vals], axis=1))
    for mode in ("0", "1"):
        os.environ["ICP_B200_STATS_MODE"] = mode
        sig, med = ctx.calc_stddevs_batch(sets)
        for k, s in enumerate(sets):
            assert np.array_equal(sig[k], oracle.calc_stddevs(s)), (mode, k)
            for a in range(2):
                assert med[k, a] == oracle.median(s[:, a])
    os.environ.pop("ICP_B200_STATS_MODE")


@pytest.mark.parametrize("dim", [2, 3])
def test_l_shape_reference_test(ctx, icp, dim):  # lib.rs:510-551 / 554-595
    true_T = oracle.transform_new(R.L_TRUE_PARAM)
    init = oracle.transform_new(np.array(R.L_TRUE_PARAM) + np.array(R.L_NOISE_PARAM))
    if dim == 2:
        src = R.L_SHAPE_2D
        dst = np.array([oracle.transform_apply(true_T, p) for p in src])
        est = icp.Icp2d.new(dst, ctx).estimate(src, icp.Transform.from_array(init), 20)
    else:
        src = R.L_SHAPE_3D
        dst = np.array([oracle.transform_xy(true_T, p) for p in src])
        est = icp.Icp3d.new(dst, ctx).estimate(src, icp.Transform.from_array(init), 20)
    for p, q in zip(src, dst):
        assert np.linalg.norm(est.transform(p[:2]) - q[:2]) < 1e-3
    o = oracle.icp_estimate(src, dst, init, 20)
    assert np.abs(est.to_array() - o["T"]).max() <= TOL     # single-pair entry: guarded warp-tree sums
    cls = icp.Icp2d if dim == 2 else icp.Icp3d
    assert np.array_equal(cls.new(dst, ctx).estimate(src, icp.Transform.from_array(init), 20, ordered=True).to_array(), o["T"])
    for ordered in (True, False):
        g = _run(ctx, icp, [src], [dst], [init], 20, dim=dim, ordered=ordered)
        _check_vs_oracle(g, [src], [dst], [init], 20, ordered=ordered)


@pytest.mark.parametrize("ordered", [True, False])
def test_bundled_scans_chain(ctx, icp, scans2d, ordered):
    """examples/scan2d.rs:62-88: src = 001 fixed, dst = 002.., warm start, 20 iterations."""
    src = scans2d[1]
    T = I6.copy()
    for k in range(2, 12):
        g = _run(ctx, icp, [src], [scans2d[k]], [T], 20, ordered=ordered)
        _check_vs_oracle(g, [src], [scans2d[k]], [T], 20, ordered=ordered)
        T = g["T"][0].copy()


@pytest.mark.parametrize("ordered", [True, False])
def test_bundled_scans_batch_with_origin_duplicates(ctx, icp, scans2d, ordered):
    ids = [2, 3, 50, 99, 100, 150, 209]     # >= 099 hold duplicated (0,0) points: index ties
    srcs = [scans2d[k - 1] for k in ids]
    dsts = [scans2d[k] for k in ids]
    T0 = [I6] * len(ids)
    g = _run(ctx, icp, srcs, dsts, T0, 20, ordered=ordered)
    _check_vs_oracle(g, srcs, dsts, T0, 20, ordered=ordered)
    # the tie is real: some query of pair 099 picked an origin duplicate, and it is the lowest index
    d = scans2d[99]
    dup = np.flatnonzero((d == 0).all(axis=1))
    assert len(dup) >= 2
    so = g["src_off"]
    c = g["corr"][int(so[3]) * 20:int(so[3]) * 20 + len(srcs[3]) * 20]
    assert dup[0] in c and not np.isin(dup[1:], c).any()


# instance ids: icp_rust_b200/csrc/Makefile (3-7 warp-tree reductions, 11-13 ordered sums)
INSTANCES = {3: (384, 3, False), 4: (256, 5, False), 5: (512, 3, False), 6: (256, 8, False), 7: (384, 8, False),
             11: (128, 9, True), 12: (256, 8, True), 13: (384, 8, True)}


@pytest.mark.parametrize("instance", sorted(INSTANCES))
def test_synthetic_pairs_every_large_instance(icp, instance):
    """config 2 shape (1080 beams, mm units, 3 mm noise, 1 % invalid returns), 20 iterations,
    on each kernel instance able to hold it."""
    os.environ["ICP_B200_FORCE_INSTANCE"] = str(instance)
    try:
        c = icp.Context(0)
    finally:
        os.environ.pop("ICP_B200_FORCE_INSTANCE")
    threads, ppt, ordered = INSTANCES[instance]
    B = 24
    src, dst, off, _ = icp.synth_scan_pairs(B, beams=1080, first_pair=1000 * instance)
    T0 = np.tile(I6, (B, 1))
    g = c.estimate_batch(src, off, dst, off, T0, 20, corr_mode=icp.CORR_ALL, ordered=ordered)
    assert c.call_info()["block_threads"] == threads   # the forced instance really ran
    g["src_off"] = off
    srcs = [src[off[b]:off[b + 1]] for b in range(B)]
    dsts = [dst[off[b]:off[b + 1]] for b in range(B)]
    _check_vs_oracle(g, srcs, dsts, list(T0), 20, ordered=ordered)
    c.close()


@pytest.mark.parametrize("ordered", [True, False])
@pytest.mark.parametrize("beams", [100, 300, 700])
def test_small_instances(ctx, icp, beams, ordered):
    B = 16
    src, dst, off, _ = icp.synth_scan_pairs(B, beams=beams, first_pair=77)
    T0 = np.tile(I6, (B, 1))
    g = ctx.estimate_batch(src, off, dst, off, T0, 8, corr_mode=icp.CORR_ALL, ordered=ordered)
    g["src_off"] = off
    _check_vs_oracle(g, [src[off[b]:off[b + 1]] for b in range(B)], [dst[off[b]:off[b + 1]] for b in range(B)],
                     list(T0), 8, ordered=ordered)


def test_ragged_and_degenerate_batch(ctx, icp, scans2d):
    rng = np.random.default_rng(3)
    a = rng.normal(0, 100, (40, 2))
    srcs = [a, a[:1], np.zeros((0, 2)), a, np.zeros((30, 2)), scans2d[1], a[:2], R.ZERO_X_SRC]
    dsts = [a.copy(), a, a, a[:1], np.zeros((30, 2)), scans2d[0], a[5:9],
            np.array([oracle.transform_apply(oracle.transform_new([0., 0.01, 0.]), p) for p in R.ZERO_X_SRC])]
    T0 = [I6] * len(srcs)
    for ordered in (True, False):
        g = _run(ctx, icp, srcs, dsts, T0, 6, ordered=ordered)
        _check_vs_oracle(g, srcs, dsts, T0, 6, nn_mode=oracle.NN_BRUTE, ordered=ordered)
        assert np.array_equal(g["T"][0], I6)                      # sigma == 0 -> None -> identity
        assert (g["reason"][0] == icp.BREAK_NONE).all()
        assert (g["reason"][2] == icp.BREAK_NONE).all()          # empty source: check_input_size


def test_status_codes_instead_of_panics(ctx, icp):
    rng = np.random.default_rng(4)
    a = rng.normal(0, 100, (20, 2))
    bad = a.copy()
    bad[3, 1] = np.nan
    srcs = [a, bad, a, a]
    dsts = [np.zeros((0, 2)), a, a, a + 0.5]
    T0 = [I6, I6, np.array([1., 0., 0., 1., np.inf, 0.]), I6]
    g = _run(ctx, icp, srcs, dsts, T0, 2, raise_on_pair_error=False)
    assert list(g["status"]) == [icp.PAIR_EMPTY_TARGET, icp.PAIR_NON_FINITE, icp.PAIR_NON_FINITE, icp.PAIR_OK]
    assert g["rc"] == icp.ERR_NON_FINITE
    for b in range(3):
        assert np.array_equal(g["T"][b], T0[b], equal_nan=True)
    with pytest.raises(icp.IcpError):
        _run(ctx, icp, srcs, dsts, T0, 2)
    # above 3072 source points the streaming kernel holds a scan, in both reduction modes
    big = rng.normal(0, 100, (3073, 2))
    g = _run(ctx, icp, [big], [big + 0.25], [I6], 1, ordered=True)
    assert ctx.call_info()["block_threads"] == 512
    _check_vs_oracle(g, [big], [big + 0.25], [I6], 1, ordered=True)
    g = _run(ctx, icp, [big], [big + 0.25], [I6], 1)
    assert ctx.call_info()["block_threads"] == 512
    _check_vs_oracle(g, [big], [big + 0.25], [I6], 1, ordered=False)
    # max_iter == 0 returns the initial transform (lib.rs:112 runs zero times)
    g = _run(ctx, icp, [a], [a + 1.0], [oracle.transform_new([1., 2., 0.3])], 0)
    assert np.array_equal(g["T"][0], oracle.transform_new([1., 2., 0.3]))


# ------------------------------------------------------------------------------------------------
@pytest.mark.parametrize("dim", [2, 3])
def test_target_search_matches_brute_force(ctx, icp, dim):  # KdTree::search, lib.rs:121,164
    rng = np.random.default_rng(11)
    m = 20000
    dst = rng.normal(0, 1000, (m, dim))
    if dim == 3:
        dst[:, 2] *= 0.05
    dst[500:520] = dst[100]           # duplicated points: lowest index must win
    dst[7000] = dst[6000]
    q = np.concatenate([rng.normal(0, 1100, (4000, dim)), dst[95:105], dst[6995:7005],
                        rng.uniform(-1e5, 1e5, (64, dim)), np.zeros((1, dim))])
    t = icp.Target(dst, ctx)
    got = t.search(q)
    exp = oracle.nn_search(dst, q, oracle.NN_BRUTE)
    assert np.array_equal(got, exp)
    assert np.array_equal(oracle.nn_search(dst, q, oracle.NN_KDTREE), exp)
    t.close()


def test_target_search_bundled_scan_ties(ctx, icp, scans2d):
    d = scans2d[120]
    t = icp.Target(d, ctx)
    q = np.concatenate([scans2d[119], np.zeros((3, 2)), [[1e-3, -1e-3]]])
    assert np.array_equal(t.search(q), oracle.nn_search(d, q, oracle.NN_BRUTE))
    t.close()


def test_shared_target_estimate(ctx, icp):
    """config 3 in small: one shared map, many scans, each against the whole map."""
    m = icp.synth_map(20000, seed=7)
    scans, off, T0 = icp.synth_map_scans(12, beams=360, seed=7)
    t = icp.Target(m, ctx)
    srcs = [scans[off[b]:off[b + 1]] for b in range(12)]
    for ordered in (True, False):
        g = t.estimate_batch(scans, off, T0, 6, corr_mode=icp.CORR_ALL, ordered=ordered)
        g["src_off"] = off
        _check_vs_oracle(g, srcs, [m] * 12, list(T0), 6, ordered=ordered)
    t.close()


# ------------------------------------------------------------------------------------------------
@pytest.mark.parametrize("ordered", [True, False])
def test_full_size_config2_properties(ctx, icp, ordered):
    """BASELINE.json config 2 at full size (4096 pairs x 1080 points x 20 iterations):
    (1) every pair agrees with the oracle (multi-threaded kd-tree run of the same batch): iteration
    counts and correspondence checksums identical, transforms bit-exact (ordered) / within 1e-9;
    (2) results do not depend on how the batch is split or ordered (same pair -> same bits);
    (3) the estimate recovers the generator's true transform to scan-noise accuracy."""
    B = 4096
    src, dst, off, true_T = icp.synth_scan_pairs(B, beams=1080)
    T0 = np.tile(I6, (B, 1))
    g = ctx.estimate_batch(src, off, dst, off, T0, 20, ordered=ordered)
    assert (g["status"] == 0).all() and np.isfinite(g["T"]).all()
    o = oracle.icp_estimate_batch(2, src, off, dst, off, T0, 20, nn_mode=oracle.NN_KDTREE)
    assert np.array_equal(g["inner"], o["inner"])
    assert np.array_equal(g["corr_hash"], o["corr_hash"])
    if ordered:
        assert np.array_equal(g["T"], o["T"])
    else:
        assert np.abs(g["T"] - o["T"]).max() <= TOL
    # split / order independence
    perm = np.random.default_rng(1).permutation(B)[:512]
    s2 = np.concatenate([src[off[b]:off[b + 1]] for b in perm])
    d2 = np.concatenate([dst[off[b]:off[b + 1]] for b in perm])
    off2 = (np.arange(513, dtype=np.uint64) * 1080).astype(np.uint32)
    g2 = ctx.estimate_batch(s2, off2, d2, off2, T0[:512], 20, ordered=ordered)
    assert np.array_equal(g2["T"], g["T"][perm]) and np.array_equal(g2["corr_hash"], g["corr_hash"][perm])
    assert np.array_equal(g2["inner"], g["inner"][perm])
    # accuracy against the generator's truth (mm / rad)
    err_t = np.abs(g["T"][:, 4:] - true_T[:, 4:]).max(axis=1)
    err_r = np.abs(g["T"][:, 2] - true_T[:, 2])
    assert np.median(err_t) < 5.0 and np.median(err_r) < 2e-3


# ------------------------------------------------------------------------------------------------
# streaming kernel (icp_stream.cuh): scans beyond the fused kernel's capacity
def test_streaming_kernel_on_config2_pairs(icp):
    """Same 1080-point pairs, forced through the large-scan kernel: correspondences and iteration
    counts identical to the oracle, transforms within 1e-9."""
    os.environ["ICP_B200_FORCE_STREAM"] = "1"
    try:
        c = icp.Context(0)
    finally:
        os.environ.pop("ICP_B200_FORCE_STREAM")
    B = 32
    src, dst, off, _ = icp.synth_scan_pairs(B, beams=1080, first_pair=4242)
    T0 = np.tile(I6, (B, 1))
    g = c.estimate_batch(src, off, dst, off, T0, 20, corr_mode=icp.CORR_ALL)
    assert c.call_info()["block_threads"] == 512
    g["src_off"] = off
    _check_vs_oracle(g, [src[off[b]:off[b + 1]] for b in range(B)], [dst[off[b]:off[b + 1]] for b in range(B)],
                     list(T0), 20, ordered=False)
    c.close()


def test_bundled_3d_frames_chain(ctx, icp, scans3d):
    """examples/scan3d.rs:104-131: src = frame 0 (27,063 points) fixed, dst = frames 0, 1, 2,
    warm start, 20 iterations, through the Icp3d facade and through the batch entry.  Frame 0
    against itself is the degenerate path (all residuals 0 -> sigma 0 -> None every iteration)."""
    src = scans3d[0]
    T = I6.copy()
    for k, dst in enumerate(scans3d):
        o = oracle.icp_estimate(src, dst, T, 20, nn_mode=oracle.NN_KDTREE, corr_mode=oracle.CORR_LAST)
        g = ctx.estimate_batch(src, np.array([0, len(src)], np.uint32), dst, np.array([0, len(dst)], np.uint32),
                               T[None], 20, dim=3, corr_mode=icp.CORR_LAST)
        assert ctx.call_info()["block_threads"] == 512
        assert np.array_equal(g["inner"][0], o["inner"])
        assert int(g["corr_hash"][0]) == int(o["corr_hash"])
        assert np.array_equal(g["corr"][:len(src)], o["corr"])
        assert np.abs(g["T"][0] - o["T"]).max() <= TOL
        est = icp.Icp3d.new(dst, ctx).estimate(src, icp.Transform.from_array(T), 20)
        assert np.abs(est.to_array() - o["T"]).max() <= TOL
        if k == 0:
            assert np.array_equal(g["T"][0], I6) and (g["reason"][0] == icp.BREAK_NONE).all()
        T = o["T"].copy()


def test_config4_shape_batch(ctx, icp, scans3d):
    """config 4 in small: pairs of 16,384-point 3-D clouds cut from the bundled frames, each pair
    moved by its own seeded SE(2) transform, seeded initial perturbation."""
    rng = np.random.default_rng(44)
    srcs, dsts, T0 = [], [], []
    for b in range(6):
        i, j = [(0, 1), (1, 2), (0, 2), (1, 0), (2, 1), (2, 0)][b]
        W = oracle.transform_new([rng.uniform(-1, 1), rng.uniform(-1, 1), rng.uniform(-3, 3)])
        def move(pts):
            xy = pts[:, :2] @ W[:4].reshape(2, 2).T + W[4:]
            return np.column_stack([xy, pts[:, 2]])
        srcs.append(move(scans3d[i][:16384]))
        dsts.append(move(scans3d[j][:16384]))
        T0.append(oracle.transform_new([rng.uniform(-0.02, 0.02), rng.uniform(-0.02, 0.02), rng.uniform(-0.01, 0.01)]))
    g = _run(ctx, icp, srcs, dsts, T0, 20, dim=3)
    assert ctx.call_info()["block_threads"] == 512
    _check_vs_oracle(g, srcs, dsts, T0, 20, ordered=False)


# ------------------------------------------------------------------------------------------------
def test_scan_chain_driver_matches_oracle_chain(ctx, icp, scans2d):
    """examples/scan2d.rs:62-105 through icp_rust_b200.scans.scan_chain: transforms bit-identical to
    the oracle's warm-started chain, path = translation of the inverse (transform.rs:26-32)."""
    from icp_rust_b200 import scans as S
    seq = [scans2d[k] for k in range(1, 8)]
    transforms, path = S.scan_chain(seq, 20, ctx)
    T = I6.copy()
    for k, est in enumerate(transforms):
        T = oracle.icp_estimate(seq[0], seq[k + 1], T, 20)["T"]
        assert np.array_equal(est.to_array(), T)
        assert np.allclose(path[k], oracle.transform_inverse(T)[4:], rtol=0, atol=1e-9)


def test_scan_chain_driver_3d(ctx, icp, scans3d):
    """examples/scan3d.rs:104-148 on the bundled frames (27k points: streaming kernel)."""
    from icp_rust_b200 import scans as S
    frames = [scans3d[0][:6000], scans3d[0][:6000], scans3d[1][:6000]]
    transforms, path = S.scan_chain(frames, 5, ctx)
    T = I6.copy()
    for k, est in enumerate(transforms):
        T = oracle.icp_estimate(frames[0], frames[k + 1], T, 5, nn_mode=oracle.NN_KDTREE)["T"]
        assert np.abs(est.to_array() - T).max() <= 1e-9
    assert path.shape == (2, 2)


def test_multi_hypothesis_front_end(ctx, icp, scans2d):
    """many initial transforms for one pair in one batch (one copy of the scan,
    ICP_B200_FLAG_SHARED_SOURCE), each scored on the device with the loop's own cost (lib.rs:45-50,
    icp_b200_batch.out_cost); the oracle's final_cost is the same quantity."""
    from icp_rust_b200 import scans as S
    src, dst = scans2d[10], scans2d[12]
    inits = [I6, oracle.transform_new([30., -20., 0.02]), oracle.transform_new([-60., 45., -0.05]),
             oracle.transform_new([400., 400., 0.6])]
    r = S.estimate_best_of(src, dst, inits, 20, ctx)
    costs = []
    for h, T0 in enumerate(inits):
        o = oracle.icp_estimate(src, dst, T0, 20, nn_mode=oracle.NN_KDTREE)
        assert np.abs(r["T"][h] - o["T"]).max() <= 1e-9
        costs.append(o["cost"])
    assert np.allclose(r["cost"], costs, rtol=1e-9, atol=0)
    assert r["best"] == int(np.argmin(costs))
    # bit-exact in ordered mode, shared source == repeated source, every kernel family
    off = np.array([0, len(src)], np.uint32)
    T0 = np.array(inits)
    t = icp.Target(dst, ctx)
    a = t.estimate_batch(src, off, T0, 20, ordered=True, want_cost=True, shared_source=True, corr_mode=icp.CORR_LAST)
    b = t.estimate_batch(np.tile(src, (4, 1)), (np.arange(5) * len(src)).astype(np.uint32), T0, 20, ordered=True,
                         want_cost=True, corr_mode=icp.CORR_LAST)
    t.close()
    for k in ("T", "inner", "corr_hash", "cost", "corr"):
        assert np.array_equal(a[k], b[k]), k
    assert np.array_equal(a["cost"], costs)
    c = ctx.estimate_batch(src, off, np.tile(dst, (4, 1)), (np.arange(5) * len(dst)).astype(np.uint32), T0, 20,
                           ordered=True, want_cost=True, shared_source=True)
    assert np.array_equal(c["T"], a["T"]) and np.array_equal(c["cost"], a["cost"])


def test_final_cost_every_kernel(ctx, icp, frames3d_20):
    """out_cost against the oracle's final huber_error: lean / fused / streaming kernels, both modes."""
    B = 12
    src, dst, off, _ = icp.synth_scan_pairs(B, beams=1080, first_pair=31337)
    T0 = np.tile(I6, (B, 1))
    o = oracle.icp_estimate_batch(2, src, off, dst, off, T0, 20, want_cost=True)
    g = ctx.estimate_batch(src, off, dst, off, T0, 20, want_cost=True)
    assert np.allclose(g["cost"], o["cost"], rtol=1e-11, atol=0)
    g = ctx.estimate_batch(src, off, dst, off, T0, 20, want_cost=True, ordered=True)
    assert np.array_equal(g["cost"], o["cost"]) and np.array_equal(g["T"], o["T"])
    from icp_rust_b200.workloads import config4_pairs
    src, dst, off, T0 = config4_pairs(3, frames=frames3d_20)
    o = oracle.icp_estimate_batch(3, src, off, dst, off, T0, 5, want_cost=True)
    g = ctx.estimate_batch(src, off, dst, off, T0, 5, dim=3, want_cost=True)
    assert np.allclose(g["cost"], o["cost"], rtol=1e-11, atol=0)
    g = ctx.estimate_batch(src, off, dst, off, T0, 5, dim=3, want_cost=True, ordered=True)   # tiled ordered chain
    assert np.array_equal(g["cost"], o["cost"]) and np.array_equal(g["T"], o["T"])
    assert np.array_equal(g["inner"], o["inner"]) and np.array_equal(g["corr_hash"], o["corr_hash"])


def test_knife_edge_guard_reruns_flagged_pairs(icp):
    """The default mode with its guard bands forced wide (ICP_B200_GUARD_SCALE): every pair is
    flagged and re-run with ordered sums inside the same call, so the results are BIT-identical to
    the oracle; with the real bands a few pairs in a thousand are re-run and every inner iteration
    count / correspondence checksum of 2048 pairs equals the oracle's; with the guard off the same
    call leaves the tree results."""
    B = 2048
    src, dst, off, _ = icp.synth_scan_pairs(B, beams=1080, first_pair=500000)
    T0 = np.tile(I6, (B, 1))
    o = oracle.icp_estimate_batch(2, src, off, dst, off, T0, 20, nn_mode=oracle.NN_KDTREE, want_margins=True)
    os.environ["ICP_B200_GUARD_SCALE"] = "1e9"
    try:
        c = icp.Context(0)
    finally:
        os.environ.pop("ICP_B200_GUARD_SCALE")
    n = 256
    g = c.estimate_batch(src[:n * 1080], off[:n + 1], dst[:n * 1080], off[:n + 1], T0[:n], 20)
    assert g["redo_pairs"] == n
    assert np.array_equal(g["T"], o["T"][:n]) and np.array_equal(g["inner"], o["inner"][:n])
    assert np.array_equal(g["corr_hash"], o["corr_hash"][:n])
    c.close()
    c = icp.Context(0)
    g = c.estimate_batch(src, off, dst, off, T0, 20)
    assert np.array_equal(g["inner"], o["inner"]) and np.array_equal(g["corr_hash"], o["corr_hash"])
    assert np.abs(g["T"] - o["T"]).max() <= TOL
    # every pair whose tightest error test lies inside the band was re-run (and is bit-identical)
    inside = np.flatnonzero(o["margins"][:, 1] < 0.5e-11)
    assert g["redo_pairs"] >= len(inside)
    assert all(np.array_equal(g["T"][b], o["T"][b]) for b in inside)
    assert g["redo_pairs"] <= B // 20, "the guard should re-run a few pairs per thousand, not a large share"
    g2 = c.estimate_batch(src, off, dst, off, T0, 20, guard=False)
    assert g2["redo_pairs"] == 0
    c.close()


def test_config3_queries_full_map(ctx, icp):
    """BASELINE config 3 at full map size: 100k-point map, the queries 400 scans issue from their
    perturbed poses (432k queries) + exact duplicates of map points; every index equals the oracle's
    exact kd-tree (lowest index on ties), through the seeded disc search of icp_target_nn.cuh."""
    m = icp.synth_map(100_000, seed=7)
    scans, off, T0 = icp.synth_map_scans(400, beams=1080, seed=7)
    q = np.concatenate([scans[off[b]:off[b + 1]] @ T0[b, :4].reshape(2, 2).T + T0[b, 4:6] for b in range(400)]
                       + [m[::500], m[::333] + 1e-7])
    t = icp.Target(m, ctx)
    got = t.search(q)
    t.close()
    exp = oracle.nn_search(m, q, oracle.NN_KDTREE)
    assert np.array_equal(got, exp)
    brute = oracle.nn_search(m, q[:3000], oracle.NN_BRUTE)
    assert np.array_equal(exp[:3000], brute)


# ------------------------------------------------------------------------------------------------
# BASELINE.json configs at their full size (configs 1, 3, 4 were sampled, not run)
@pytest.mark.parametrize("ordered", [True, False])
def test_config1_whole_chain_and_batch(ctx, icp, scans2d, ordered):
    """BASELINE config 1 in full: examples/scan2d.rs:62-88 — src = 001 fixed, dst = 002 ... 209,
    warm start, 20 iterations — every frame of the recording, (a) as the chain the example runs and
    (b) as one 208-pair batch from identity."""
    src = scans2d[1]
    ids = list(range(2, 210))
    # (a) the chain.  Each frame starts from the ORACLE's previous transform so that every one of
    # the 208 estimates is compared on identical inputs; the device's own chain is carried beside it.
    T_or, T_dev = I6.copy(), I6.copy()
    so = np.array([0, len(src)], np.uint32)
    for k in ids:
        dst = scans2d[k]
        do = np.array([0, len(dst)], np.uint32)
        o = oracle.icp_estimate(src, dst, T_or, 20, nn_mode=oracle.NN_KDTREE)
        g = ctx.estimate_batch(src, so, dst, do, T_or[None], 20, ordered=ordered)
        assert np.array_equal(g["inner"][0], o["inner"]), f"frame {k}: inner iteration counts differ"
        assert int(g["corr_hash"][0]) == int(o["corr_hash"]), f"frame {k}: correspondences differ"
        if ordered:
            assert np.array_equal(g["T"][0], o["T"]), f"frame {k}"
        else:
            assert np.abs(g["T"][0] - o["T"]).max() <= TOL, f"frame {k}"
        T_dev = ctx.estimate_batch(src, so, dst, do, T_dev[None], 20, ordered=ordered, want_inner=False,
                                   want_hash=False)["T"][0].copy()
        T_or = o["T"].copy()
        assert np.abs(T_dev - T_or).max() <= TOL, f"frame {k}: the device's own chain drifted"
    if ordered:
        assert np.array_equal(T_dev, T_or)
    # (b) the same 208 pairs as one batch from identity
    srcs = [src] * len(ids)
    dsts = [scans2d[k] for k in ids]
    s, so = _cat(srcs, 2)
    d, do = _cat(dsts, 2)
    T0 = np.tile(I6, (len(ids), 1))
    g = ctx.estimate_batch(s, so, d, do, T0, 20, ordered=ordered)
    o = oracle.icp_estimate_batch(2, s, so, d, do, T0, 20, nn_mode=oracle.NN_KDTREE)
    assert np.array_equal(g["inner"], o["inner"]) and np.array_equal(g["corr_hash"], o["corr_hash"])
    assert (g["status"] == 0).all()
    if ordered:
        assert np.array_equal(g["T"], o["T"])
    else:
        assert np.abs(g["T"] - o["T"]).max() <= TOL


def test_config3_icp_loop_full_map(ctx, icp):
    """BASELINE config 3 at full map size inside the ICP loop: 1024 scans x 1080 points, each against
    the WHOLE shared 100,000-point map, 20 iterations; oracle = one kd-tree (Icp2d::new once), one
    estimate per scan.  Default mode on all 1024, ordered sums on the first 128."""
    m = icp.synth_map(100_000, seed=7)
    B = 1024
    scans, off, T0 = icp.synth_map_scans(B, beams=1080, seed=7)
    o = oracle.icp_estimate_batch(2, scans, off, m, None, T0, 20, nn_mode=oracle.NN_KDTREE)
    t = icp.Target(m, ctx)
    g = t.estimate_batch(scans, off, T0, 20)
    assert (g["status"] == 0).all()
    assert np.array_equal(g["inner"], o["inner"])
    assert np.array_equal(g["corr_hash"], o["corr_hash"])
    assert np.abs(g["T"] - o["T"]).max() <= TOL
    n = 128
    g = t.estimate_batch(scans[:n * 1080], off[:n + 1], T0[:n], 20, ordered=True)
    assert np.array_equal(g["inner"], o["inner"][:n]) and np.array_equal(g["corr_hash"], o["corr_hash"][:n])
    assert np.array_equal(g["T"], o["T"][:n])
    t.close()


def test_config4_baseline_shape(ctx, icp, frames3d_20):
    """BASELINE config 4 at its own shape: pairs (frame i, frame i+1 mod 20) of the bundled 3-D
    recording (examples/scan3d.rs:104-131), 16,384 points per cloud, seeded initial perturbation,
    20 iterations; 80 pairs = every frame pair four times with different initial transforms."""
    from icp_rust_b200.workloads import config4_pairs
    B = 80
    src, dst, off, T0 = config4_pairs(B, frames=frames3d_20)
    o = oracle.icp_estimate_batch(3, src, off, dst, off, T0, 20, nn_mode=oracle.NN_KDTREE)
    g = ctx.estimate_batch(src, off, dst, off, T0, 20, dim=3)
    assert ctx.call_info()["block_threads"] == 512
    assert (g["status"] == 0).all()
    assert np.array_equal(g["inner"], o["inner"])
    assert np.array_equal(g["corr_hash"], o["corr_hash"])
    assert np.abs(g["T"] - o["T"]).max() <= TOL


def test_target_update_matches_fresh_build(ctx, icp):
    """icp_b200_target_update: after a series of removals and additions — inside the
    bounding box (incremental path: geometry and seeds kept) and outside it (rebuild) — every search
    equals a freshly built target's and the oracle's on the same point set; estimates too."""
    rng = np.random.default_rng(21)
    m = icp.synth_map(30000, seed=7)
    pts = m.copy()
    t = icp.Target(pts, ctx)
    scans, off, T0 = icp.synth_map_scans(6, beams=1080, seed=7)
    q = np.concatenate([scans[off[b]:off[b + 1]] @ T0[b, :4].reshape(2, 2).T + T0[b, 4:6] for b in range(6)]
                       + [rng.uniform(m.min(0), m.max(0), (500, 2))])
    for step in range(5):
        rem = rng.choice(len(pts), size=int(rng.integers(1, 3000)), replace=False).astype(np.uint32)
        if step == 1:
            rem = np.concatenate([rem, rem[:50]])          # duplicates in remove_idx are tolerated
        if step == 2:
            rem = np.arange(len(pts) - 400, len(pts), dtype=np.uint32)   # a removed tail
        base = pts[rng.integers(0, len(pts), 2500)]
        add = base + rng.normal(0, 15.0, base.shape)
        lo, hi = m.min(0), m.max(0)
        add = np.clip(add, lo, hi)                          # inside the box of the build: incremental path
        if step == 3:
            add = np.concatenate([add, [[hi[0] + 5000.0, hi[1] + 100.0]]])   # outside: the structure is rebuilt
        t.update(add, rem)
        keep = np.ones(len(pts), bool)
        keep[rem] = False
        pts = np.concatenate([pts[keep], add])
        assert t.m == len(pts)
        got = t.search(q)
        fresh = icp.Target(pts, ctx)
        assert np.array_equal(got, fresh.search(q)), f"step {step}"
        fresh.close()
        assert np.array_equal(got, oracle.nn_search(pts, q, oracle.NN_KDTREE)), f"step {step}"
    n = 4
    g = t.estimate_batch(scans[:n * 1080], off[:n + 1], T0[:n], 10)
    o = oracle.icp_estimate_batch(2, scans[:n * 1080], off[:n + 1], pts, None, T0[:n], 10)
    assert np.array_equal(g["inner"], o["inner"]) and np.array_equal(g["corr_hash"], o["corr_hash"])
    # 3-D target, everything removed, then refilled
    p3 = rng.normal(0, 1, (2000, 3))
    t3 = icp.Target(p3, ctx)
    t3.update(None, np.arange(2000, dtype=np.uint32))
    assert t3.m == 0
    t3.update(p3[:700] * 0.5, None)
    q3 = rng.normal(0, 1, (300, 3))
    assert np.array_equal(t3.search(q3), oracle.nn_search(p3[:700] * 0.5, q3, oracle.NN_BRUTE))
    t3.close()
    t.close()


@pytest.mark.parametrize("forced_reruns", [False, True])
def test_device_pointer_calls_back_to_back(icp, forced_reruns):
    """icp_b200_estimate_batch_device, asynchronous, with ICP_B200_FLAG_INDEPENDENT: calls of very
    different sizes queued on one stream without anything between them (each may start while the
    one before it drains) give, bit for bit, what the same batches give one at a time through the
    host-pointer entry; with the guard bands forced wide every pair also goes through the re-run
    launch of its call while the next call is already resident.  The library's per-call state
    (two parities of counters, workspaces and flagged-pair lists) is reused every other call."""
    import torch
    sizes = [600, 37, 1, 900, 5, 444, 2, 148]
    total = sum(sizes)
    src, dst, off, _ = icp.synth_scan_pairs(total, beams=1080, first_pair=740000)
    if forced_reruns:
        os.environ["ICP_B200_GUARD_SCALE"] = "1e9"
    try:
        c = icp.Context(0)
    finally:
        os.environ.pop("ICP_B200_GUARD_SCALE", None)
    dev = torch.device("cuda:0")
    stream = torch.cuda.Stream(device=dev)
    want, bufs, first = [], [], 0
    for n in sizes:
        s = src[first * 1080:(first + n) * 1080]
        d = dst[first * 1080:(first + n) * 1080]
        o = (off[first:first + n + 1] - off[first]).astype(np.uint32)
        T0 = np.tile(I6, (n, 1))
        want.append(c.estimate_batch(s, o, d, o, T0, 20))
        if forced_reruns:
            assert want[-1]["redo_pairs"] == n
        bufs.append(dict(src=torch.from_numpy(s.copy()).to(dev), dst=torch.from_numpy(d.copy()).to(dev),
                         off=torch.from_numpy(o.astype(np.int32)).to(dev), T0=torch.from_numpy(T0).to(dev),
                         out=torch.zeros((n, 6), dtype=torch.float64, device=dev),
                         inner=torch.zeros((n, 20), dtype=torch.int32, device=dev),
                         hash=torch.zeros(n, dtype=torch.int64, device=dev),
                         status=torch.full((n,), -1, dtype=torch.int32, device=dev)))
        first += n
    torch.cuda.synchronize()

    def enqueue(k, flags, synchronize=False):
        b = bufs[k]
        c.estimate_batch_device(dict(src=b["src"].data_ptr(), src_off=b["off"].data_ptr(), dst=b["dst"].data_ptr(),
                                     dst_off=b["off"].data_ptr(), init_T=b["T0"].data_ptr(), out_T=b["out"].data_ptr(),
                                     out_inner=b["inner"].data_ptr(), out_corr_hash=b["hash"].data_ptr(),
                                     out_status=b["status"].data_ptr()),
                                sizes[k], 20, dim=2, max_src=1080, max_dst=1080, stream=stream.cuda_stream,
                                synchronize=synchronize, flags=flags)

    for rounds, flags in ((3, icp.FLAG_INDEPENDENT), (1, 0)):
        for b in bufs:
            b["out"].zero_()
            b["status"].fill_(-1)
        torch.cuda.synchronize()
        with torch.cuda.stream(stream):
            for _ in range(rounds):
                for k in range(len(sizes)):
                    enqueue(k, flags)
        torch.cuda.synchronize()
        for k, b in enumerate(bufs):
            assert np.array_equal(b["out"].cpu().numpy(), want[k]["T"]), f"call {k} ({sizes[k]} pairs), flags {flags}"
            assert np.array_equal(b["inner"].cpu().numpy().astype(np.uint32), want[k]["inner"])
            assert np.array_equal(b["hash"].cpu().numpy().astype(np.uint64), want[k]["corr_hash"])
            assert int(b["status"].abs().sum().item()) == 0
    # a synchronous call after the asynchronous ones reports its own re-run count
    enqueue(3, 0, synchronize=True)
    info = c.call_info()
    assert info["redo_pairs"] == (sizes[3] if forced_reruns else want[3]["redo_pairs"])
    c.close()


def test_outer_fixed_point_exit_is_exact(icp, scans2d, frames3d_20):
    """icp_fused.cuh fused_repeats: outer iterations that provably repeat (an inner solve accepted
    no update, so the outer transform stays) are written, not recomputed.  A context that computes
    every iteration (ICP_B200_FIXED_POINT=0) gives the same bits in every output — transforms,
    inner-iteration words, correspondences of all 20 iterations, checksums, final costs — on the
    bundled 2-D scans (lean and ordered kernels), the config-4 clouds (streaming kernel) and scans
    against a shared map; and on all three the exit is actually taken."""
    from icp_rust_b200.workloads import config4_pairs
    os.environ["ICP_B200_FIXED_POINT"] = "0"
    try:
        full = icp.Context(0)
    finally:
        os.environ.pop("ICP_B200_FIXED_POINT")
    fast = icp.Context(0)
    idx = list(range(0, 120, 3))
    s2, so = _cat([scans2d[k] for k in idx], 2)
    d2, do = _cat([scans2d[k + 1] for k in idx], 2)
    T2 = np.tile(I6, (len(idx), 1))
    s3, d3, o3, T3 = config4_pairs(6, frames=frames3d_20)
    m = icp.synth_map(20000, seed=7)
    sm, om, Tm = icp.synth_map_scans(24, beams=360, seed=7)
    keys = ("T", "inner", "corr", "corr_hash", "cost", "status")
    maps = {fast: icp.Target(m, fast), full: icp.Target(m, full)}
    taken = {}
    for ordered in (False, True):
        for name, call in (
                ("2d", lambda c: c.estimate_batch(s2, so, d2, do, T2, 20, corr_mode=icp.CORR_ALL, want_cost=True, ordered=ordered)),
                ("3d", lambda c: c.estimate_batch(s3, o3, d3, o3, T3, 20, dim=3, corr_mode=icp.CORR_ALL, want_cost=True, ordered=ordered)),
                ("map", lambda c: maps[c].estimate_batch(sm, om, Tm, 20, corr_mode=icp.CORR_ALL, want_cost=True, ordered=ordered))):
            a, b = call(fast), call(full)
            for k in keys:
                assert np.array_equal(a[k], b[k]), (name, ordered, k)
            idle = (a["inner"] & 0xffff) == 0
            taken[name] = float((idle[:, :-1].any(axis=1)).mean())
    assert taken["2d"] > 0.5 and taken["3d"] > 0.5, taken
    # correspondences of the LAST iteration only
    a = fast.estimate_batch(s2, so, d2, do, T2, 20, corr_mode=icp.CORR_LAST)
    b = full.estimate_batch(s2, so, d2, do, T2, 20, corr_mode=icp.CORR_LAST)
    assert np.array_equal(a["corr"], b["corr"])
    for t in maps.values():
        t.close()
    fast.close()
    full.close()


def test_overlapping_calls_other_kernels(icp, frames3d_20):
    """ICP_B200_FLAG_INDEPENDENT through the other kernels: asynchronous device-pointer calls on the
    streaming kernel (3-D clouds beyond the fused instances), with ordered sums, and against a
    shared target (icp_b200_target_estimate_batch_device) — queued back to back, twice over, the
    results equal those of the synchronous host-pointer calls bit for bit.  max_iter 1 and 2: the
    fixed-point exit at the edge of the loop."""
    import torch
    from icp_rust_b200.workloads import config4_init_transforms
    c = icp.Context(0)
    dev = torch.device("cuda:0")
    stream = torch.cuda.Stream(device=dev)
    n3 = 3600
    s3, o3 = _cat([frames3d_20[k + 1][:n3] for k in range(3)], 3)
    d3, _ = _cat([frames3d_20[k][:n3] for k in range(3)], 3)
    T3 = config4_init_transforms(3)
    src, dst, off, _ = icp.synth_scan_pairs(40, beams=1080, first_pair=31000)
    T2 = np.tile(I6, (40, 1))
    m = icp.synth_map(20000, seed=7)
    sm, om, Tm = icp.synth_map_scans(30, beams=360, seed=7)
    tgt = icp.Target(m, c)

    def to_dev(a, dtype=None):
        t = torch.from_numpy(np.ascontiguousarray(a))
        return (t.to(dtype) if dtype else t).to(dev)

    cases = []
    for max_iter in (1, 2, 20):
        cases.append(dict(name=f"stream3d/{max_iter}", dim=3, B=3, n=n3, max_iter=max_iter, ordered=False, target=None,
                          want=c.estimate_batch(s3, o3, d3, o3, T3, max_iter, dim=3),
                          bufs=dict(src=to_dev(s3), dst=to_dev(d3), off=to_dev(o3.astype(np.int32)), T0=to_dev(T3))))
    cases.append(dict(name="ordered2d", dim=2, B=40, n=1080, max_iter=20, ordered=True, target=None,
                      want=c.estimate_batch(src, off, dst, off, T2, 20, ordered=True),
                      bufs=dict(src=to_dev(src), dst=to_dev(dst), off=to_dev(off.astype(np.int32)), T0=to_dev(T2))))
    cases.append(dict(name="shared_map", dim=2, B=30, n=360, max_iter=20, ordered=False, target=tgt,
                      want=tgt.estimate_batch(sm, om, Tm, 20),
                      bufs=dict(src=to_dev(sm), dst=None, off=to_dev(om.astype(np.int32)), T0=to_dev(Tm))))
    for cs in cases:
        cs["out"] = torch.zeros((cs["B"], 6), dtype=torch.float64, device=dev)
        cs["hash"] = torch.zeros(cs["B"], dtype=torch.int64, device=dev)
    torch.cuda.synchronize()
    with torch.cuda.stream(stream):
        for _ in range(2):
            for cs in cases:
                b = cs["bufs"]
                ptrs = dict(src=b["src"].data_ptr(), src_off=b["off"].data_ptr(), init_T=b["T0"].data_ptr(),
                            out_T=cs["out"].data_ptr(), out_corr_hash=cs["hash"].data_ptr())
                if cs["target"] is None:
                    ptrs.update(dst=b["dst"].data_ptr(), dst_off=b["off"].data_ptr())
                c.estimate_batch_device(ptrs, cs["B"], cs["max_iter"], dim=cs["dim"], max_src=cs["n"], max_dst=cs["n"],
                                        stream=stream.cuda_stream, synchronize=False, target=cs["target"], ordered=cs["ordered"],
                                        flags=icp.FLAG_INDEPENDENT)
    torch.cuda.synchronize()
    for cs in cases:
        assert np.array_equal(cs["out"].cpu().numpy(), cs["want"]["T"]), cs["name"]
        assert np.array_equal(cs["hash"].cpu().numpy().astype(np.uint64), cs["want"]["corr_hash"]), cs["name"]
    tgt.close()
    c.close()


@pytest.mark.gpu
@pytest.mark.parametrize("forced_reruns", [False, True])
def test_submit_wait_pipeline_one_thread(icp, forced_reruns):
    """icp_b200_estimate_batch_submit / _wait: one thread streams batches of very different sizes
    through two contexts, batch k+1 enqueued before batch k is waited for (pinned buffers).  Every
    batch gives, bit for bit, what the one-call form gives; a context with a submitted batch refuses
    other compute calls; wait without a submitted batch is a no-op; a rejected pair surfaces at wait.
    With the guard bands forced wide every pair goes through the re-run that wait issues."""
    import torch
    sizes = [300, 5, 444, 1, 150, 37]
    total = sum(sizes)
    src, dst, off, _ = icp.synth_scan_pairs(total, beams=1080, first_pair=910000)
    if forced_reruns:
        os.environ["ICP_B200_GUARD_SCALE"] = "1e9"
    try:
        cs = [icp.Context(0), icp.Context(0)]
    finally:
        os.environ.pop("ICP_B200_GUARD_SCALE", None)
    assert cs[0].wait() == 0
    want, bufs, first = [], [], 0
    for n in sizes:
        s = src[first * 1080:(first + n) * 1080]
        d = dst[first * 1080:(first + n) * 1080]
        o = (off[first:first + n + 1] - off[first]).astype(np.uint32)
        T0 = np.tile(I6, (n, 1))
        want.append(cs[0].estimate_batch(s, o, d, o, T0, 20))
        if forced_reruns:
            assert want[-1]["redo_pairs"] == n
        pin = lambda a: torch.from_numpy(np.ascontiguousarray(a)).pin_memory()
        bufs.append(dict(src=pin(s), dst=pin(d), off=pin(o.astype(np.int32)), T0=pin(T0),
                         out=torch.zeros((n, 6), dtype=torch.float64).pin_memory(),
                         inner=torch.zeros((n, 20), dtype=torch.int32).pin_memory(),
                         hash=torch.zeros(n, dtype=torch.int64).pin_memory(),
                         status=torch.full((n,), -1, dtype=torch.int32).pin_memory()))
        first += n

    def ptrs(b):
        return dict(src=b["src"].data_ptr(), src_off=b["off"].data_ptr(), dst=b["dst"].data_ptr(),
                    dst_off=b["off"].data_ptr(), init_T=b["T0"].data_ptr(), out_T=b["out"].data_ptr(),
                    out_inner=b["inner"].data_ptr(), out_corr_hash=b["hash"].data_ptr(), out_status=b["status"].data_ptr())

    cs[0].submit_host_ptrs(ptrs(bufs[0]), sizes[0], 20)
    # the context is taken until its wait
    with pytest.raises(icp.IcpError):
        cs[0].submit_host_ptrs(ptrs(bufs[1]), sizes[1], 20)
    with pytest.raises(icp.IcpError):
        cs[0].estimate_batch_host_ptrs(ptrs(bufs[1]), sizes[1], 20)
    for k in range(1, len(sizes)):
        cs[k % 2].submit_host_ptrs(ptrs(bufs[k]), sizes[k], 20)
        cs[(k - 1) % 2].wait()
        if forced_reruns:
            assert cs[(k - 1) % 2].call_info()["redo_pairs"] == sizes[k - 1]
    cs[(len(sizes) - 1) % 2].wait()
    for k, n in enumerate(sizes):
        b, w = bufs[k], want[k]
        assert np.array_equal(b["out"].numpy(), w["T"]), k
        assert np.array_equal(b["inner"].numpy().view(np.uint32), w["inner"]), k
        assert np.array_equal(b["hash"].numpy().view(np.uint64), w["corr_hash"]), k
        assert (b["status"].numpy() == 0).all()

    # a rejected pair (NaN coordinate) is reported by wait, as the one-call form reports it
    bad = bufs[1]
    bad["src"][3, 0] = float("nan")
    cs[0].submit_host_ptrs(ptrs(bad), sizes[1], 20)
    assert cs[0].wait(allow=(icp.ERR_NON_FINITE,)) == icp.ERR_NON_FINITE
    assert bad["status"].numpy()[0] == 3 and (bad["status"].numpy()[1:] == 0).all()
    assert cs[0].wait() == 0
    # shared target: submit on the target, wait on its context
    m = icp.synth_map(20000, seed=11)
    scans, soff, sT0 = icp.synth_map_scans(8, beams=360, seed=11)
    tgt = icp.Target(m, cs[1])
    w = tgt.estimate_batch(scans, soff, sT0, 10)
    out = torch.zeros((8, 6), dtype=torch.float64).pin_memory()
    sc, so, st = (torch.from_numpy(np.ascontiguousarray(a)).pin_memory() for a in (scans, soff.astype(np.int32), sT0))
    cs[1].submit_host_ptrs(dict(src=sc.data_ptr(), src_off=so.data_ptr(), init_T=st.data_ptr(), out_T=out.data_ptr()),
                           8, 10, target=tgt)
    cs[1].wait()
    assert np.array_equal(out.numpy(), w["T"])
    tgt.close()
    for c in cs:
        c.close()
