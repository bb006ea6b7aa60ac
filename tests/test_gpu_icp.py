"""GPU parity: ICP path (A1-A3) through the C ABI against the reference goldens and the CPU
oracle.  Tolerance: T within 1e-9 absolute of the float64 reference (the north star allows
1e-5 relative); iteration counts equal."""
import numpy as np
import pytest

from conftest import icp_cases, load_golden

pytestmark = pytest.mark.gpu

torch = pytest.importorskip("torch")

T_ATOL = 1e-9


@pytest.fixture(scope="module")
def env():
    import b2slam
    from b2slam import _lib, devapi, synth
    from oracle import corc
    if _lib.device_count() <= 0:
        pytest.fail("no CUDA device: the gpu tests must run on the B200 box")

    class E:
        pass
    e = E()
    e.b2slam, e.lib, e.dev, e.synth, e.corc = b2slam, _lib, devapi, synth, corc
    e.icp = b2slam.ICP()
    return e


@pytest.mark.parametrize("case", icp_cases(), ids=lambda c: "seed%d_n%d_it%d" % (
    int(c["seed"]), c["src"].shape[1], int(c["max_iter"])))
def test_process_matches_reference_golden(env, case):
    icp = env.b2slam.ICP(max_iter=int(case["max_iter"]), tolerance=float(case["tol"]))
    tar = env.synth.homogeneous(case["tar"].astype(np.float64))
    src = env.synth.homogeneous(case["src"].astype(np.float64))
    keep_t, keep_s = tar.copy(), src.copy()
    T = icp.process(tar, src)                      # target first, like the reference
    assert T.shape == (3, 3) and T.dtype == np.float64
    assert icp.last_iterations == int(case["iters"])
    np.testing.assert_allclose(T, case["T"], rtol=0, atol=T_ATOL)
    assert np.array_equal(T[2], [0.0, 0.0, 1.0])
    assert np.array_equal(tar, keep_t) and np.array_equal(src, keep_s)  # inputs untouched


def test_find_nearest_ties_and_golden(env):
    z = load_golden("icp_pieces.npz")
    for s, t, d, i in (("tie_src", "tie_tar", "tie_dist", "tie_idx"),
                       ("sym_src", "sym_tar", "sym_dist", "sym_idx")):
        dist, idx = env.icp.findNearest(z[s], z[t])
        assert np.array_equal(idx, z[i])
        np.testing.assert_allclose(dist, z[d], rtol=0, atol=1e-14)
    big_s = np.random.Generator(np.random.PCG64(5)).uniform(-10, 10, (3000, 2))
    big_t = np.random.Generator(np.random.PCG64(6)).uniform(-10, 10, (2500, 2))
    dist, idx = env.icp.findNearest(big_s, big_t)
    od, oi = env.corc.nearest(big_s, big_t)
    assert np.array_equal(idx, oi) and np.allclose(dist, od, rtol=0, atol=1e-14)


def test_find_nearest_sqrt_ties_follow_the_reference(env):
    """src (0,0), tar [(1, 2^-26), (-1, 0)] -> both norms are exactly 1.0, the reference returns
    index 0, an argmin over squared distances returns 1.  Plus the 10^4 generated cases of the golden."""
    z = load_golden("icp_ties.npz")
    dist, idx = env.icp.findNearest(z["repro_src"], z["repro_tar"])
    assert int(idx[0]) == 0 and dist[0] == 1.0
    # every case is its own (1 x 6) problem; run them as one batch of block-diagonal problems by offsetting the clouds
    # far apart (exact: the offsets are multiples of 2^10, coordinates below 2^4, so sums and differences are exact
    # in float64 up to 2^-39 ... NOT exact) -> instead call case by case for a slice and in bulk through process below
    for c in range(0, 10000, 23):
        d, i = env.icp.findNearest(z["gen_src"][c:c + 1], z["gen_tar"][c])
        assert int(i[0]) == int(z["gen_idx"][c]) and d[0] == z["gen_dist"][c], c


def test_batched_search_resolves_sqrt_ties_like_the_reference(env):
    """The batched kernel (every search mode) on clouds whose first-iteration correspondences are all decided by the
    tie rule: T and the iteration count of the unmodified reference (max_iter 1 and the defaults)."""
    z = load_golden("icp_ties.npz")
    L = env.lib.lib()
    try:
        for prune in (0, 1, 2, 3, 4):
            assert L.b2s_tune(b"icp_prune", prune) == 0
            for n in range(int(z["icp_count"])):
                icp = env.b2slam.ICP(max_iter=int(z["icp%d_max_iter" % n]), tolerance=float(z["icp%d_tol" % n]))
                tar, src = z["icp%d_tar" % n], z["icp%d_src" % n]
                T, it = icp.process_batch(np.repeat(tar[None], 3, 0), np.repeat(src[None], 3, 0))
                assert (it == int(z["icp%d_iters" % n])).all(), (prune, n)
                np.testing.assert_allclose(T[1], z["icp%d_T" % n], rtol=0, atol=T_ATOL, err_msg="prune %d case %d" % (prune, n))
    finally:
        L.b2s_tune(b"icp_prune", 4)
    # the 10^4 generated single-point cases through the batched kernel with max_iter = 1: the transform of a one-point
    # source is the pure translation onto its match, so T's translation names the chosen target
    tar = np.ascontiguousarray(np.transpose(z["gen_tar"], (0, 2, 1)))            # (P, 2, 6)
    src = np.ascontiguousarray(z["gen_src"][:, :, None])                          # (P, 2, 1)
    icp = env.b2slam.ICP(max_iter=1, tolerance=0.0)
    T, it = icp.process_batch(tar, src)
    want = z["gen_tar"][np.arange(len(tar)), z["gen_idx"]] - z["gen_src"]
    others = z["gen_tar"] - z["gen_src"][:, None, :]
    got_idx = np.abs(others - T[:, None, :2, 2]).sum(-1).argmin(1)
    assert np.array_equal(got_idx, z["gen_idx"]), "%d cases matched another target" % int((got_idx != z["gen_idx"]).sum())
    np.testing.assert_allclose(T[:, :2, 2], want, rtol=0, atol=1e-12)


def test_get_transform_golden_including_reflections(env):
    z = load_golden("icp_pieces.npz")
    for a, b, T in zip(z["fit_src"], z["fit_tar"], z["fit_T"]):
        np.testing.assert_allclose(env.icp.getTransform(a, b), T, rtol=0, atol=1e-12)
    with pytest.raises(ValueError):
        env.icp.getTransform(np.zeros((4, 2)), np.zeros((5, 2)))


@pytest.mark.parametrize("beams,pairs", [(120, 96), (360, 64), (1080, 12)])
def test_batch_vs_oracle(env, beams, pairs):
    tar, src, _ = env.synth.icp_pairs(4001, pairs, beams)
    want_T, want_it = env.corc.icp_batch(tar, src, 30, 1e-3)
    T32, it32 = env.icp.process_batch(tar, src)                                   # float32 inputs
    T64, it64 = env.icp.process_batch(tar.astype(np.float64), src.astype(np.float64))
    bad = int((it32 != want_it).sum())
    assert bad == 0, "%d pairs stopped at a different iteration" % bad
    assert np.array_equal(it32, it64)
    np.testing.assert_allclose(T32, want_T, rtol=0, atol=T_ATOL)
    np.testing.assert_allclose(T64, T32, rtol=0, atol=1e-13)


def test_w7_launch_parameters_run_exactly_max_iter(env):
    tar, src, _ = env.synth.icp_pairs(7001, 1, 360)
    T, iters = env.icp.process_batch(tar, src, max_iter=10, tolerance=0.0)
    assert iters[0] == 10                      # strict '<': tolerance 0 never breaks
    want_T, _ = env.corc.icp_batch(tar, src, 10, 0.0)
    np.testing.assert_allclose(T, want_T, rtol=0, atol=T_ATOL)


def test_unequal_sizes_odd_counts_and_zero_iterations(env):
    tar, src, _ = env.synth.icp_pairs(91, 5, 150)
    src = np.ascontiguousarray(src[:, :, :97])       # N=97 (odd: no bulk-copy alignment), M=150
    want_T, want_it = env.corc.icp_batch(tar, src, 30, 1e-3)
    T, it = env.icp.process_batch(tar, src)
    assert np.array_equal(it, want_it)
    np.testing.assert_allclose(T, want_T, rtol=0, atol=T_ATOL)
    tar2 = np.ascontiguousarray(tar[:, :, :77])      # M=77: unaligned target rows
    want_T, want_it = env.corc.icp_batch(tar2, src, 30, 1e-3)
    T, it = env.icp.process_batch(tar2, src)
    assert np.array_equal(it, want_it)
    np.testing.assert_allclose(T, want_T, rtol=0, atol=T_ATOL)
    T0, it0 = env.icp.process_batch(tar, src, max_iter=0)
    assert (it0 == 0).all()
    np.testing.assert_allclose(T0, np.broadcast_to(np.identity(3), T0.shape), rtol=0, atol=1e-12)
    Te, ite = env.icp.process_batch(tar[:0], src[:0])
    assert Te.shape == (0, 3, 3) and ite.shape == (0,)


def test_single_pair_graph_path(env):
    """ICP.process (one pair) replays a captured CUDA graph; it must give the bits of the plain stream path and of
    the batched call, survive changing sizes / parameters / dtypes (re-capture) and interleaved batched calls
    (device buffers may move), and keep returning fresh results on replay."""
    tune = env.lib.lib().b2s_tune
    rng = np.random.Generator(np.random.PCG64(5))
    seen = []
    try:
        for rep, (n, m, dtype, max_iter, tol) in enumerate([(360, 360, np.float32, 30, 1e-3), (360, 360, np.float32, 30, 1e-3),
                                                          (120, 97, np.float64, 30, 1e-3), (360, 360, np.float32, 10, 0.0),
                                                          (1080, 1080, np.float32, 30, 1e-3), (360, 360, np.float32, 30, 1e-3)]):
            tar, src, _ = env.synth.icp_pairs(100 + rep, 1, max(n, m))
            tar, src = tar[:, :, :m].astype(dtype), src[:, :, :n].astype(dtype)
            assert tune(b"icp_graph", 1) == 0
            Tg, ig = env.icp.process_batch(tar, src, max_iter=max_iter, tolerance=tol)
            Tg2, ig2 = env.icp.process_batch(tar, src, max_iter=max_iter, tolerance=tol)      # replay
            assert tune(b"icp_graph", 0) == 0
            Tp, ip = env.icp.process_batch(tar, src, max_iter=max_iter, tolerance=tol)
            assert np.array_equal(Tg, Tp) and np.array_equal(ig, ip) and np.array_equal(Tg2, Tp), rep
            both = np.concatenate([tar, tar]), np.concatenate([src, src])
            Tb, ib = env.icp.process_batch(*both, max_iter=max_iter, tolerance=tol)          # batched call in between
            assert np.array_equal(Tb[0], Tp[0]) and np.array_equal(Tb[1], Tp[0]) and ib[0] == ip[0]
            seen.append(Tg)
        assert not np.array_equal(seen[0], seen[1])      # different pairs, same shape: the replay read the new inputs
        assert tune(b"icp_graph", 1) == 0
        cloud = np.vstack([rng.normal(0, 3, (2, 200)), np.ones((1, 200))])
        T = env.icp.process(cloud, cloud)                # the reference-signature call
        np.testing.assert_allclose(T, np.identity(3), rtol=0, atol=1e-12)
    finally:
        tune(b"icp_graph", 1)


def test_call_timeline_trace_does_not_change_results(env, capfd):
    """B2S_TRACE=1 prints the timeline of the host-buffer calls (and routes single pairs through the plain stream
    path); results are unchanged and the marks arrive on stderr."""
    import os
    xy, _ = env.synth.room_sequence(9001, 40, 360)
    T0, it0 = env.icp.process_sequence(xy)
    ox, oy, cx, cy = env.synth.grid_scans(3, 6, 360, half_extent_m=5.0)
    m0 = env.b2slam.Mapping(256, 256, 0.05)
    p0 = m0.update_batch(ox, oy, cx, cy).copy()
    os.environ["B2S_TRACE"] = "1"
    try:
        T1, it1 = env.icp.process_sequence(xy)
        T2, it2 = env.icp.process_batch(xy[:1], xy[1:2])
        m1 = env.b2slam.Mapping(256, 256, 0.05)
        p1 = m1.update_batch(ox, oy, cx, cy).copy()
    finally:
        del os.environ["B2S_TRACE"]
    err = capfd.readouterr().err
    assert "[b2s trace]" in err and "icp chunk done" in err and "ray-cast chunk done" in err
    assert np.array_equal(T1, T0) and np.array_equal(it1, it0) and np.array_equal(T2[0], T0[0])
    assert np.array_equal(p1, p0)


def test_largest_supported_scan_and_the_limit(env):
    """2304 points per scan is the documented ceiling (the CTA's register budget): it must run and agree with the
    oracle; one point more is refused with a status code, not a launch failure."""
    tar, src, _ = env.synth.icp_pairs(2400, 2, 2304)
    T, it = env.icp.process_batch(tar, src)
    want_T, want_it = env.corc.icp_batch(tar, src, 30, 1e-3)
    assert np.array_equal(it, want_it)
    np.testing.assert_allclose(T, want_T, rtol=0, atol=T_ATOL)
    big = np.zeros((1, 2, 2305), dtype=np.float32)
    with pytest.raises(Exception) as err:
        env.icp.process_batch(big, big)
    assert "2304" in str(err.value)


def test_device_pointer_abi_and_sequence_chain(env):
    """Layer 1 on CUDA tensors; cfg-2 style consecutive pairs of a room sequence."""
    xy, _ = env.synth.room_sequence(9001, 65, 360)
    tar = torch.from_numpy(xy[:-1]).cuda().contiguous()
    src = torch.from_numpy(xy[1:]).cuda().contiguous()
    T, it = env.dev.icp_batch(tar, src, 30, 1e-3)
    torch.cuda.synchronize()
    want_T, want_it = env.corc.icp_batch(xy[:-1], xy[1:], 30, 1e-3)
    assert np.array_equal(it.cpu().numpy(), want_it)
    np.testing.assert_allclose(T.cpu().numpy(), want_T, rtol=0, atol=T_ATOL)
    from b2slam import scan
    from oracle import pyref
    traj = scan.compose_odometry((0.0, 0.0, 0.0), T.cpu().numpy())
    st = (0.0, 0.0, 0.0)
    for t in want_T:
        st = pyref.compose_pose(st, t)
    assert np.allclose(traj[-1], st, rtol=0, atol=1e-7)


def test_sequence_form_equals_pair_form(env):
    """process_sequence == process_batch on the consecutive pairs, bit for bit (float32 and float64 input),
    including streams too short to hold a pair and chunk boundaries of the copy pipeline."""
    xy, _ = env.synth.room_sequence(9001, 300, 360)
    for arr in (xy, xy.astype(np.float64)):
        T, it = env.icp.process_sequence(arr)
        T2, it2 = env.icp.process_batch(arr[:-1], arr[1:])
        assert T.shape == (299, 3, 3) and np.array_equal(T, T2) and np.array_equal(it, it2)
    tune = env.lib.lib().b2s_tune
    try:
        for chunks in (1, 3, 8):
            assert tune(b"h2d_chunks", chunks) == 0
            T3, it3 = env.icp.process_sequence(xy[:50])
            assert np.array_equal(T3, T2[:49]) and np.array_equal(it3, it2[:49])
    finally:
        tune(b"h2d_chunks", 0)
    want_T, want_it = env.corc.icp_batch(xy[:8], xy[1:9], 30, 1e-3)
    assert np.array_equal(it[:8], want_it)
    np.testing.assert_allclose(T[:8], want_T, rtol=0, atol=T_ATOL)
    for k in (0, 1):
        Te, ite = env.icp.process_sequence(xy[:k])
        assert Te.shape == (0, 3, 3) and ite.shape == (0,)
    T1, _ = env.icp.process_sequence(xy[:2])
    assert np.array_equal(T1[0], T2[0])


def test_odometry_is_sequence_plus_pose_chain(env):
    """ICP.odometry == process_sequence + the sequential pose loop of [ICP]:185-190 on its transforms."""
    from oracle import pyref
    xy, _ = env.synth.room_sequence(9001, 400, 360)
    traj, T, it = env.icp.odometry(xy, state=(0.5, -0.25, 0.1))
    T2, it2 = env.icp.process_sequence(xy)
    assert np.array_equal(T, T2) and np.array_equal(it, it2)
    st = (0.5, -0.25, 0.1)
    want = [st]
    for t in T:
        st = pyref.compose_pose(st, t)
        want.append(st)
    np.testing.assert_allclose(traj, np.array(want), rtol=0, atol=1e-10)
    one, T1, it1 = env.icp.odometry(xy[:1], state=(1.0, 2.0, 3.0))
    assert one.shape == (1, 3) and np.array_equal(one[0], [1.0, 2.0, 3.0]) and T1.shape == (0, 3, 3) and it1.shape == (0,)


def test_process_scans_equals_laser_to_numpy_plus_sequence(env):
    """Fused ingestion for the ICP: ranges in, laserToNumpy inside the kernel.  Bit-identical to the
    host conversion ([ICP]:216-229 / the W12 clamp [SLAM]:115-123) followed by process_sequence, and equal to the
    oracle on the converted clouds."""
    import math
    from b2slam import scan
    rng = np.random.Generator(np.random.PCG64(31))
    for beams, scans in ((360, 120), (1080, 24), (100, 9), (361, 6)):   # 361: rows not 16-byte multiples
        xy, _ = env.synth.room_sequence(9100 + beams, scans, beams)
        ranges = np.hypot(xy[:, 0, :], xy[:, 1, :]).astype(np.float32)
        for clamp in (None, 30.0):
            r = ranges.copy()
            if clamp is not None:
                r[rng.integers(0, scans, 5), rng.integers(0, beams, 5)] = np.inf
            clouds = np.stack([scan.laser_to_points(r[k], -math.pi, math.pi, clamp_inf_to=clamp)[:2] for k in range(scans)])
            want_T, want_it = env.icp.process_sequence(clouds)
            T, it = env.icp.process_scans(r, -math.pi, math.pi, clamp_inf_to=clamp)
            assert np.array_equal(it, want_it) and np.array_equal(T, want_T), (beams, clamp)
            traj, T2, it2 = env.icp.process_scans(r, -math.pi, math.pi, clamp_inf_to=clamp, state=(1.0, -2.0, 0.3))
            traj3, T3, it3 = env.icp.odometry(clouds, state=(1.0, -2.0, 0.3))
            assert np.array_equal(T2, T) and np.array_equal(traj, traj3)
        oT, oit = env.corc.icp_batch(clouds[:-1], clouds[1:], 30, 1e-3)
        assert np.array_equal(it, oit)
        np.testing.assert_allclose(T, oT, rtol=0, atol=T_ATOL)
    Te, ite = env.icp.process_scans(np.zeros((1, 16), dtype=np.float32), -1.0, 1.0)
    assert Te.shape == (0, 3, 3) and ite.shape == (0,)


def test_cfg2_full_sequence_properties(env):
    """cfg 2 at full size (10 000 scans, 9 999 pairs): every T is a proper rigid transform, a scan matched onto
    itself gives the identity in one iteration ([ICP]:75-77: mean error 0 -> break), and 512 evenly spaced pairs
    agree with the oracle."""
    xy, _ = env.synth.room_sequence(9001, 10000, 360)
    T, it = env.icp.process_sequence(xy)
    assert T.shape == (9999, 3, 3) and it.min() >= 1 and it.max() <= 30
    R = T[:, :2, :2]
    assert np.abs(np.einsum("pij,pkj->pik", R, R) - np.identity(2)).max() < 1e-12
    assert np.abs(np.linalg.det(R) - 1.0).max() < 1e-12
    assert np.array_equal(T[:, 2, :], np.broadcast_to([0.0, 0.0, 1.0], (9999, 3)))
    pick = np.linspace(0, 9998, 512).astype(int)
    want_T, want_it = env.corc.icp_batch(xy[pick], xy[pick + 1], 30, 1e-3)
    assert np.array_equal(it[pick], want_it)
    np.testing.assert_allclose(T[pick], want_T, rtol=0, atol=T_ATOL)
    Ts, its = env.icp.process_batch(xy[:256], xy[:256])
    assert (its == 1).all()
    np.testing.assert_allclose(Ts, np.broadcast_to(np.identity(3), Ts.shape), rtol=0, atol=1e-12)


def test_cfg4_shape_sample_vs_oracle(env):
    """cfg 4 shape (1080-beam independent pairs) at a size the oracle finishes in seconds, plus a
    size-independent check on a larger batch: every pair of a batch gives the result it gives alone."""
    tar, src, _ = env.synth.icp_pairs(4001, 2048, 1080)
    T, it = env.icp.process_batch(tar, src)
    pick = np.linspace(0, 2047, 24).astype(int)
    want_T, want_it = env.corc.icp_batch(tar[pick], src[pick], 30, 1e-3)
    assert np.array_equal(it[pick], want_it)
    np.testing.assert_allclose(T[pick], want_T, rtol=0, atol=T_ATOL)
    T2, it2 = env.icp.process_batch(tar[1000:1100], src[1000:1100])     # batch position must not matter
    assert np.array_equal(it2, it[1000:1100]) and np.array_equal(T2, T[1000:1100])
    # every search mode (0 brute force, 1 per-lane block pruning, 2 warp-level + per-lane, 3 warp-level only) and
    # every pruning block size gives the brute-force answer bit for bit
    tune = env.lib.lib().b2s_tune
    try:
        for prune, block in ((0, 0), (1, 16), (1, 32), (2, 8), (2, 16), (2, 32), (3, 8), (3, 16), (4, 8), (4, 16)):
            assert tune(b"icp_prune", prune) == 0 and tune(b"icp_block", block) == 0
            T3, it3 = env.icp.process_batch(tar[:256], src[:256])
            assert np.array_equal(it3, it[:256]) and np.array_equal(T3, T[:256]), (prune, block)
    finally:
        tune(b"icp_prune", 4)
        tune(b"icp_block", 0)


def test_non_finite_target_points_are_never_matched(env):
    """[ICP]:99-106: a NaN / inf distance never satisfies `dist < min_dist`, so such targets are ignored and the
    solve stays finite; the pruned searches must not skip anything because of them (same bits as brute force)."""
    tar, src, _ = env.synth.icp_pairs(4242, 6, 360)
    tar = tar.copy()
    tar[0, 0, 5] = np.nan
    tar[1, 1, 200] = np.inf
    tar[2, :, 17] = np.nan
    tar[3, 0, 0] = -np.inf
    tar[4, :, 100:140] = np.nan
    want_T, want_it = env.corc.icp_batch(tar, src, 30, 1e-3)
    assert np.isfinite(want_T).all()
    tune = env.lib.lib().b2s_tune
    got = {}
    try:
        for prune in (0, 1, 2, 3, 4):
            assert tune(b"icp_prune", prune) == 0
            got[prune] = env.icp.process_batch(tar, src)
    finally:
        tune(b"icp_prune", 4)
    for prune in (0, 1, 2, 3, 4):
        T, it = got[prune]
        assert np.array_equal(it, want_it), prune
        np.testing.assert_allclose(T, want_T, rtol=0, atol=T_ATOL)
        assert np.array_equal(T, got[0][0])


# ----------------------------------------------------------------------------- adjacent steps

def test_queued_search_overflow_paths_equal_brute_force(env):
    """The queued search (icp_prune 4) hands a group of 32 points to the collective search when a point needs more than
    three blocks or the group has more than 16 candidate blocks, takes a second level of bounds above 64 blocks, and falls
    back to 16-target blocks above 254 blocks.  Shapes that force each of those, float32 and float64 clouds, ragged last
    blocks: T and the iteration counts equal the brute force bit for bit, and the oracle within tolerance."""
    rng = np.random.Generator(np.random.PCG64(4242))
    cases = []
    # (a) a tight cluster: every block is within reach of every point -> every group overflows its queue
    tar = rng.normal(0, 0.02, (3, 2, 333)); src = tar[:, :, rng.integers(0, 333, 301)] + rng.normal(0, 0.005, (3, 2, 301))
    cases.append(("cluster", tar, src))
    # (b) a coarse first bound: the source is the target rotated by 0.5 rad, so the first iterations need many blocks
    ang = np.linspace(-np.pi, np.pi, 700); r = 5 + np.sin(3 * ang)
    tar = np.stack([r * np.cos(ang), r * np.sin(ang)])[None].repeat(2, 0) + rng.normal(0, 0.01, (2, 2, 700))
    c, s_ = np.cos(0.5), np.sin(0.5)
    src = np.stack([c * tar[:, 0] - s_ * tar[:, 1], s_ * tar[:, 0] + c * tar[:, 1]], axis=1)[:, :, ::-1][:, :, :650].copy()
    cases.append(("rotated", tar, src))
    # (c) 1501 targets: 188 blocks of 8 -> second level of the warp test, ragged last block; (d) 2301: 16-target blocks
    for m, n in ((1501, 1490), (2301, 2304)):
        ang = np.linspace(-np.pi, np.pi, m); r = 6 + 2 * np.sin(2 * ang + 0.3)
        tar = np.stack([r * np.cos(ang), r * np.sin(ang)])[None] + rng.normal(0, 0.01, (1, 2, m))
        pick = np.sort(rng.integers(0, m, n))
        c, s_ = np.cos(0.03), np.sin(0.03)
        src = np.stack([c * tar[:, 0, pick] - s_ * tar[:, 1, pick] + 0.05, s_ * tar[:, 0, pick] + c * tar[:, 1, pick] - 0.04], axis=1)
        cases.append(("large%d" % m, tar, src + rng.normal(0, 0.01, src.shape)))
    tune = env.lib.lib().b2s_tune
    for name, tar, src in cases:
        for dtype in (np.float32, np.float64):
            t, s = np.ascontiguousarray(tar.astype(dtype)), np.ascontiguousarray(src.astype(dtype))
            got = {}
            try:
                for prune in ((0, 4) if name.startswith("large") or dtype is np.float32 else (2, 4)):
                    assert tune(b"icp_prune", prune) == 0
                    got[prune] = env.icp.process_batch(t, s)
            finally:
                tune(b"icp_prune", 4)
            (Ta, ia), (Tb, ib) = got.values()
            assert np.array_equal(ia, ib) and np.array_equal(Ta, Tb), (name, dtype)
            if t.shape[2] <= 800:
                want_T, want_it = env.corc.icp_batch(t, s, 30, 1e-3)
                assert np.array_equal(ib, want_it), (name, dtype)
                np.testing.assert_allclose(Tb, want_T, rtol=0, atol=T_ATOL, err_msg=name)


def test_pose_chain_parallel_prefix_vs_sequential_loop(env):
    from b2slam import scan
    from oracle import pyref
    xy, _ = env.synth.room_sequence(9001, 2001, 360)
    T, _ = env.icp.process_batch(xy[:-1], xy[1:])
    traj = scan.compose_odometry_gpu((0.5, -0.25, 0.1), T)
    st = (0.5, -0.25, 0.1)
    want = [st]
    for t in T:
        st = pyref.compose_pose(st, t)
        want.append(st)
    np.testing.assert_allclose(traj, np.array(want), rtol=0, atol=1e-10)
    assert np.allclose(traj, scan.compose_odometry((0.5, -0.25, 0.1), T), rtol=0, atol=1e-10)
    one = scan.compose_odometry_gpu((1.0, 2.0, 3.0), np.zeros((0, 3, 3)))
    assert one.shape == (1, 3) and np.array_equal(one[0], [1.0, 2.0, 3.0])


def test_virtual_scan_vs_reference_loop(env):
    import math
    from b2slam import scan
    from oracle import pyref
    rng = np.random.Generator(np.random.PCG64(17))
    # obstacle cells of a 129 x 129 map at 0.155 m (course_agv_gazebo/config/map.yaml), walls + clutter
    cells = np.argwhere((rng.random((129, 129)) < 0.03) | (np.arange(129)[:, None] % 128 == 0) | (np.arange(129)[None, :] % 128 == 0))
    obstacle = np.vstack((cells[:, 0] * 0.155 - 10.0, cells[:, 1] * 0.155 - 10.0))
    beams = 120
    inc = 2 * math.pi / (beams - 1)
    for pose in ((0.0, 0.0, 0.0), (3.2, -4.1, 1.3), (-7.7, 8.8, -2.9)):
        got = scan.virtual_scan(obstacle, pose, -math.pi, inc, beams)
        want = pyref.virtual_scan(obstacle, pose, -math.pi, inc, beams)
        # same bins, same minima; CUDA's hypot differs from libm's by at most a couple of ulp
        np.testing.assert_allclose(got, want, rtol=1e-14, atol=0)
    empty = scan.virtual_scan(np.zeros((2, 0)), (0, 0, 0), -math.pi, inc, beams)
    assert (empty == 100.0).all()


# ----------------------------------------------------------------------------- randomised sweep

def test_random_icp_shapes_vs_oracle(env):
    """Seeded fuzz over cloud sizes (down to one point), duplicated targets, collinear and far-offset clouds,
    all search modes and pruning block sizes: identical iteration counts, T within 1e-9 (relative to the cloud scale)."""
    rng = np.random.Generator(np.random.PCG64(777))
    for trial in range(30):
        n = int(rng.choice([1, 2, 3, 5, 16, 17, 31, 33, 64, 100, 127, 250]))
        m = int(rng.choice([1, 2, 4, 15, 16, 17, 32, 48, 100, 129, 300]))
        pairs = int(rng.integers(1, 5))
        scale = float(rng.choice([0.5, 5.0, 40.0]))
        tar = rng.normal(0, scale, (pairs, 2, m))
        if trial % 3 == 0:                      # collinear wall
            tar[:, 1, :] = 0.3 * tar[:, 0, :] + 1.0
        if m > 3 and trial % 4 == 0:            # duplicated targets: lowest index must win
            tar[:, :, m // 2] = tar[:, :, 0]
            tar[:, :, m - 1] = tar[:, :, 0]
        pick = rng.integers(0, m, (pairs, n))
        src = np.take_along_axis(tar, np.broadcast_to(pick[:, None, :], (pairs, 2, n)), axis=2)
        th = rng.uniform(-0.1, 0.1)
        c, s = np.cos(th), np.sin(th)
        src = np.stack([c * src[:, 0] - s * src[:, 1] + rng.uniform(-0.2, 0.2),
                        s * src[:, 0] + c * src[:, 1] + rng.uniform(-0.2, 0.2)], axis=1)
        src = src + rng.normal(0, 0.01 * scale, src.shape)
        tar = np.ascontiguousarray(tar.astype(np.float32))
        src = np.ascontiguousarray(src.astype(np.float32))
        want_T, want_it = env.corc.icp_batch(tar, src, 30, 1e-3)
        for prune in (4, 3, 2, 1, 0):
            assert env.lib.lib().b2s_tune(b"icp_prune", prune) == 0
            assert env.lib.lib().b2s_tune(b"icp_block", (0, 8, 16, 32)[trial % 4]) == 0
            try:
                T, it = env.icp.process_batch(tar, src)
            finally:
                env.lib.lib().b2s_tune(b"icp_prune", 4)
                env.lib.lib().b2s_tune(b"icp_block", 0)
            assert np.array_equal(it, want_it), "trial %d n %d m %d prune %d" % (trial, n, m, prune)
            np.testing.assert_allclose(T, want_T, rtol=0, atol=1e-9 * max(1.0, scale),
                                       err_msg="trial %d n %d m %d prune %d" % (trial, n, m, prune))


def test_pose_chain_matches_reference_publishResult_golden(env):
    """f-2 pinned: the device prefix scan against sensor_sta of the reference's own publishResult after every one of
    400 transforms (tests/golden/next_rows.npz).  The parallel prefix re-associates the sums: 1e-12 absolute."""
    from b2slam import scan
    z = load_golden("next_rows.npz")
    traj = scan.compose_odometry_gpu(tuple(z["chain_start"]), z["chain_T"])
    np.testing.assert_allclose(traj, z["chain_traj"], rtol=0, atol=1e-12)
    assert np.array_equal(traj[0], z["chain_start"])


def test_virtual_scan_matches_reference_laserEstimation_golden(env):
    """f-4 pinned: same bearing bins as the reference's own laserEstimation, ranges to 1e-13 relative (hypot on the
    device vs libm), incl. an empty map and obstacles beyond the 100.0 the empty bins hold."""
    from b2slam import scan
    z = load_golden("next_rows.npz")
    for i in range(int(z["vscan_count"])):
        g = lambda k: z["vscan%d_%s" % (i, k)]
        got = scan.virtual_scan(np.stack([g("obs_x"), g("obs_y")]), g("pose"), float(g("angle_min")),
                                float(g("angle_increment")), int(g("beams")))
        want = g("ranges")
        assert np.array_equal(got == 100.0, want == 100.0), i       # the same bins were hit
        np.testing.assert_allclose(got, want, rtol=1e-13, atol=0)


def test_streamed_sequence_calls_equal_blocking_calls(env):
    """ICP.submit_sequence / submit_scans + IcpTicket.wait (two in flight) give exactly what process_sequence /
    process_scans give, call by call, also when the streams differ in length."""
    import math
    xy, _ = env.synth.room_sequence(515, 700, 360)
    chunks = [xy[0:300], xy[250:700], xy[100:101], xy[400:520]]
    want = [env.icp.process_sequence(c) for c in chunks]
    tickets, got = [], []
    for k, c in enumerate(chunks):
        tickets.append(env.icp.submit_sequence(c))
        if k >= 1:
            T, it = tickets[k - 1].wait()
            got.append((T.copy(), it.copy()))
    T, it = tickets[-1].wait()
    got.append((T.copy(), it.copy()))
    for (T, it), (wT, wit) in zip(got, want):
        assert np.array_equal(T, wT) and np.array_equal(it, wit)
    rng = np.hypot(xy[:200, 0], xy[:200, 1]).astype(np.float32)
    wT, wit = env.icp.process_scans(rng, -math.pi, math.pi)
    t1 = env.icp.submit_scans(rng, -math.pi, math.pi)
    t2 = env.icp.submit_scans(rng[:50], -math.pi, math.pi)
    T, it = t1.wait()
    assert np.array_equal(T, wT) and np.array_equal(it, wit)
    T2, it2 = t2.wait()
    assert np.array_equal(T2, wT[:49]) and np.array_equal(it2, wit[:49])
