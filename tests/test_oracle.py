"""The CPU oracle (oracle/pyref.py literal form, oracle/oracle.c compiled form) against the
golden vectors that oracle/make_golden.py produced by executing the reference classes."""
import numpy as np
import pytest

from conftest import icp_cases, load_golden
from oracle import corc, pyref
import b2slam.synth as synth


# ----------------------------------------------------------------------------- Bresenham (A7)

def _golden_paths():
    z = load_golden("bresenham.npz")
    return z["segs"], z["offsets"], z["cells"]


def test_bresenham_pyref_matches_reference_paths():
    segs, offs, cells = _golden_paths()
    step = 3  # pure-Python loop: every third segment keeps the CPU suite short
    for i in range(0, len(segs), step):
        x0, y0, x1, y1 = (int(v) for v in segs[i])
        want = [tuple(r) for r in cells[offs[i]:offs[i + 1]].tolist()]
        assert pyref.bresenham_cells([x0, y0], [x1, y1]) == want, segs[i]


def test_bresenham_c_matches_reference_paths():
    segs, offs, cells = _golden_paths()
    for i in range(len(segs)):
        x0, y0, x1, y1 = (int(v) for v in segs[i])
        want = [tuple(r) for r in cells[offs[i]:offs[i + 1]].tolist()]
        assert corc.bresenham([x0, y0], [x1, y1]) == want, segs[i]


def test_bresenham_invariants():
    rng = np.random.Generator(np.random.PCG64(1))
    for _ in range(2000):
        a = rng.integers(-300, 300, size=2)
        b = rng.integers(-300, 300, size=2)
        p = corc.bresenham(a, b)
        if (a == b).all():
            assert p == []
            continue
        assert p[0] == tuple(a) and p[-1] == tuple(b)
        assert len(p) == max(abs(int(b[0] - a[0])), abs(int(b[1] - a[1]))) + 1
        assert p == corc.bresenham(b, a)[::-1]  # canonical trace direction
        d = np.abs(np.diff(np.array(p), axis=0))
        assert d.max() <= 1


def test_bresenham_is_not_the_integer_algorithm():
    """The float64 accumulator disagrees with exact-arithmetic Bresenham on some slopes;
    the oracle must reproduce the reference there."""
    def exact(dx, dy):
        return [(k, (2 * k * dy + dx) // (2 * dx)) for k in range(dx + 1)]
    differing = [(dx, dy) for dx in range(1, 60) for dy in range(0, dx + 1)
                 if corc.bresenham([0, 0], [dx, dy]) != exact(dx, dy)]
    assert (10, 3) in differing or (10, 7) in differing or len(differing) > 50


# ----------------------------------------------------------------------------- ICP pieces (A2, A3)

def test_nearest_ties_lowest_index_wins():
    z = load_golden("icp_pieces.npz")
    for s, t, d, i in (("tie_src", "tie_tar", "tie_dist", "tie_idx"),
                       ("sym_src", "sym_tar", "sym_dist", "sym_idx")):
        for fn in (pyref.nearest_targets, pyref.nearest_targets_vec, corc.nearest):
            dist, idx = fn(z[s], z[t])
            assert np.array_equal(np.asarray(idx, dtype=np.int64), z[i]), fn
            np.testing.assert_allclose(dist, z[d], rtol=0, atol=1e-15)
    assert z["tie_idx"][0] == 3  # triplicated target 3/17/30 -> index 3


def test_nearest_sqrt_ties_follow_the_reference():
    """Ties that exist only after the reference's sqrt (np.linalg.norm): the squared distances differ in the last
    bits, the norms are equal, the lowest index wins ([ICP]:102-103).  Golden: the reference on 10^4 such cases."""
    z = load_golden("icp_ties.npz")
    for fn in (pyref.nearest_targets, corc.nearest):
        dist, idx = fn(z["repro_src"], z["repro_tar"])
        assert int(idx[0]) == 0 == int(z["repro_idx"][0]) and dist[0] == 1.0
    assert int(z["gen_d2_argmin_differs"]) > 100          # an argmin over squared distances is NOT the reference
    got = np.array([corc.nearest(s[None], t)[1][0] for s, t in zip(z["gen_src"], z["gen_tar"])])
    assert np.array_equal(got, z["gen_idx"])
    gd = np.array([corc.nearest(s[None], t)[0][0] for s, t in zip(z["gen_src"][:500], z["gen_tar"][:500])])
    assert np.array_equal(gd, z["gen_dist"][:500])
    for c in range(0, 400):                                # literal port on a slice (pure Python)
        assert int(pyref.nearest_targets(z["gen_src"][c:c + 1], z["gen_tar"][c])[1][0]) == int(z["gen_idx"][c])


def test_icp_process_on_tie_decided_clouds_matches_reference():
    z = load_golden("icp_ties.npz")
    for n in range(int(z["icp_count"])):
        tar, src = z["icp%d_tar" % n], z["icp%d_src" % n]
        T, iters = corc.icp_batch(tar[None], src[None], int(z["icp%d_max_iter" % n]), float(z["icp%d_tol" % n]))
        assert int(iters[0]) == int(z["icp%d_iters" % n])
        np.testing.assert_allclose(T[0], z["icp%d_T" % n], rtol=0, atol=1e-12)


def test_rigid_fit_matches_reference_including_reflection_branch():
    z = load_golden("icp_pieces.npz")
    for a, b, T in zip(z["fit_src"], z["fit_tar"], z["fit_T"]):
        np.testing.assert_allclose(pyref.rigid_fit_svd(a, b), T, rtol=0, atol=1e-13)
        np.testing.assert_allclose(pyref.rigid_fit_closed_form(a, b), T, rtol=0, atol=1e-12)
        np.testing.assert_allclose(corc.rigid_fit(a, b), T, rtol=0, atol=1e-12)
        assert abs(np.linalg.det(T[:2, :2]) - 1.0) < 1e-12  # always a proper rotation


# ----------------------------------------------------------------------------- ICP.process (A1)

@pytest.mark.parametrize("case", icp_cases(), ids=lambda c: "seed%d_n%d_it%d" % (
    int(c["seed"]), c["src"].shape[1], int(c["max_iter"])))
def test_icp_process_c_matches_reference(case):
    T, iters = corc.icp_batch(case["tar"][None], case["src"][None], int(case["max_iter"]),
                              float(case["tol"]))
    assert int(iters[0]) == int(case["iters"])
    np.testing.assert_allclose(T[0], case["T"], rtol=0, atol=1e-12)


def test_icp_process_pyref_matches_reference_small():
    cases = [c for c in icp_cases() if c["src"].shape[1] <= 120][:3]
    for c in cases:
        T, iters = pyref.icp_process(synth.homogeneous(c["tar"].astype(np.float64)),
                                     synth.homogeneous(c["src"].astype(np.float64)),
                                     int(c["max_iter"]), float(c["tol"]))
        assert iters == int(c["iters"])
        np.testing.assert_allclose(T, c["T"], rtol=0, atol=1e-13)


def test_icp_recovers_known_motion():
    tar, src, truth = synth.icp_pairs(11, 4, 360)
    T, iters = corc.icp_batch(tar, src, 30, 1e-3)
    # T maps the source scan into the target frame: it must shrink the mean NN distance
    # (point-to-point ICP with the 1e-3 stop rule only approaches the inverse motion)
    for p in range(4):
        s2 = src[p].astype(np.float64).T
        t2 = tar[p].astype(np.float64).T
        before = corc.nearest(s2, t2)[0].mean()
        moved = s2 @ T[p, :2, :2].T + T[p, :2, 2]
        after = corc.nearest(moved, t2)[0].mean()
        assert after < 0.7 * before
        assert iters[p] >= 2


# ----------------------------------------------------------------------------- Mapping (A4-A6)

@pytest.mark.parametrize("tag,w_hit", [("w20", 20.0), ("w4", 4.0)])
def test_mapping_counts_reproduce_reference_maps(tag, w_hit):
    z = load_golden("mapping.npz")
    S, Hx, Hy = pyref.grid_scale(200, 200, 0.1)
    assert (S, Hx, Hy) == (10.0, 10.0, 10.0)  # the literals of mapping.py:33-36
    hit = np.zeros((200, 200), dtype=np.int32)
    miss = np.zeros((200, 200), dtype=np.int32)
    hit_py = np.zeros((200, 200), dtype=np.int32)
    miss_py = np.zeros((200, 200), dtype=np.int32)
    visits = 0
    for ox, oy, cx, cy in zip(z[tag + "_ox"], z[tag + "_oy"], z[tag + "_cx"], z[tag + "_cy"]):
        visits += corc.grid_raycast(hit, miss, S, Hx, Hy, ox[None], oy[None], [cx], [cy])
        pyref.grid_update_counts(hit_py, miss_py, ox.astype(np.float64), oy.astype(np.float64),
                                 float(cx), float(cy), S, Hx, Hy)
    assert np.array_equal(hit, hit_py) and np.array_equal(miss, miss_py)
    assert visits == int(hit.sum() + miss.sum())
    score, pmap = corc.grid_finalize(hit, miss, w_hit)
    score_py, pmap_py = pyref.finalize_counts(hit, miss, w_hit)
    assert np.array_equal(pmap, pmap_py)
    np.testing.assert_allclose(score, z[tag + "_datamap"], rtol=1e-12, atol=0)
    ambiguous = pyref.boundary_ambiguous(hit, miss, w_hit)
    ok = (pmap == z[tag + "_pmap"]) | ambiguous
    assert ok.all()
    if w_hit == 20.0:
        assert not ambiguous.any() and np.array_equal(pmap, z[tag + "_pmap"])


def _score_from_counts(hit, miss, cells, w_hit):
    return 0.01 * miss.reshape(-1)[cells].astype(np.float64) + w_hit * hit.reshape(-1)[cells].astype(np.float64)


@pytest.mark.parametrize("tag", ["g200", "g4096", "g16384"])
def test_mapping_float64_inputs_reproduce_reference(tag):
    """The reference fed UNROUNDED float64 endpoints / sensor positions (what slam_ekf.py:89-90 passes), at three
    grid scales, incl. endpoints exactly on decimal cell boundaries: the float64 oracle touches exactly the
    reference's cells with the reference's evidence; narrowing the inputs to float32 first does not."""
    z = load_golden("mapping_f64.npz")
    side = int(z[tag + "_side"])
    hit = np.zeros((side, side), dtype=np.int32)
    miss = np.zeros((side, side), dtype=np.int32)
    corc.grid_raycast(hit, miss, 10.0, 10.0, 10.0, z[tag + "_ox"], z[tag + "_oy"], z[tag + "_cx"], z[tag + "_cy"])
    cells = z[tag + "_cells"]
    touched = np.flatnonzero((hit.reshape(-1) != 0) | (miss.reshape(-1) != 0))
    assert np.array_equal(touched, cells)
    np.testing.assert_allclose(_score_from_counts(hit, miss, cells, 20.0), z[tag + "_score"], rtol=1e-12, atol=0)
    h, m = hit.reshape(-1)[cells], miss.reshape(-1)[cells]
    pm = np.where(0.01 * m + 20.0 * h > 10.0, 100, 0).astype(np.int8)
    assert np.array_equal(pm, z[tag + "_pmap"])
    if tag == "g200":
        cw = z[tag + "_cells_w4"]
        assert np.array_equal(cw, cells)
        np.testing.assert_allclose(_score_from_counts(hit, miss, cw, 4.0), z[tag + "_score_w4"], rtol=1e-12, atol=0)
        # the literal port agrees scan by scan on the small map
        h2 = np.zeros((side, side), dtype=np.int32)
        m2 = np.zeros((side, side), dtype=np.int32)
        for ox, oy, cx, cy in zip(z[tag + "_ox"], z[tag + "_oy"], z[tag + "_cx"], z[tag + "_cy"]):
            pyref.grid_update_counts(h2, m2, ox, oy, float(cx), float(cy), 10.0, 10.0, 10.0)
        assert np.array_equal(h2, hit) and np.array_equal(m2, miss)
    # float32 narrowing is NOT the reference: these very scans land in other cells
    assert int(z[tag + "_moved_by_f32"]) > 0
    h32 = np.zeros((side, side), dtype=np.int32)
    m32 = np.zeros((side, side), dtype=np.int32)
    fin = np.where(np.isfinite(z[tag + "_ox"]), z[tag + "_ox"], 0.0)
    corc.grid_raycast(h32, m32, 10.0, 10.0, 10.0, fin.astype(np.float32), z[tag + "_oy"].astype(np.float32),
                      z[tag + "_cx"].astype(np.float32), z[tag + "_cy"].astype(np.float32))
    assert not (np.array_equal(h32, hit) and np.array_equal(m32, miss))


def test_miss_stream_threshold_matches_reference():
    z = load_golden("mapping.npz")
    for m, score, pm in zip(z["miss_stream_counts"], z["miss_stream_score"], z["miss_stream_pmap"]):
        hit = np.zeros((1, 1), dtype=np.int32)
        miss = np.full((1, 1), int(m), dtype=np.int32)
        s, p = corc.grid_finalize(hit, miss)
        assert int(p[0, 0]) == int(pm)            # 1000 traversals free, 1001 occupied
        assert abs(s[0, 0] - score) < 1e-9


def test_evidence_form_equals_reference_datamap():
    z = load_golden("mapping.npz")
    dm = np.zeros((200, 200))
    pm = 50 * np.ones((200, 200))
    for ox, oy, cx, cy in zip(z["w20_ox"][:3], z["w20_oy"][:3], z["w20_cx"], z["w20_cy"]):
        pyref.grid_update_evidence(dm, pm, ox.astype(np.float64), oy.astype(np.float64),
                                   float(cx), float(cy), 10.0, 10.0, 10.0)
    hit = np.zeros((200, 200), dtype=np.int32)
    miss = np.zeros((200, 200), dtype=np.int32)
    corc.grid_raycast(hit, miss, 10.0, 10.0, 10.0, z["w20_ox"][:3], z["w20_oy"][:3],
                      z["w20_cx"][:3], z["w20_cy"][:3])
    _, pmap = corc.grid_finalize(hit, miss)
    assert np.array_equal(pmap, pm.astype(np.int8))


def test_grid_generalised_scale():
    assert pyref.grid_scale(4096, 4096, 0.05) == (20.0, 102.4, 102.4)
    assert pyref.grid_scale(16384, 16384, 0.05) == (20.0, 409.6, 409.6)
    assert pyref.world_to_cell(-10.05, 10.0, 10.0) == 0  # truncation toward zero, not floor


def test_grid_rejects_nan():
    hit = np.zeros((8, 8), dtype=np.int32)
    miss = np.zeros((8, 8), dtype=np.int32)
    with pytest.raises(ValueError):
        corc.grid_raycast(hit, miss, 1.0, 4.0, 4.0, [[np.nan]], [[0.0]], [0.0], [0.0])


# ----------------------------------------------------------------------------- scan ingestion (A8 + u2T)

def test_ingestion_oracle_matches_reference_node():
    """laserToNumpy + u2T(xEst).dot(np_msg) + Mapping.update as run by the reference node (golden)."""
    import math
    import b2slam.scan as scan
    z = load_golden("ingestion.npz")
    K, N = z["ranges"].shape
    for k in range(K):
        ox, oy = pyref.scan_to_world(z["ranges"][k], z["poses"][k], float(z["angle_min"]), float(z["angle_max"]))
        assert np.array_equal(ox, z["ox"][k]) and np.array_equal(oy, z["oy"][k])   # literal form: bit-equal
    hit = np.zeros((200, 200), dtype=np.int32)
    miss = np.zeros((200, 200), dtype=np.int32)
    corc.grid_raycast_ranges(hit, miss, 10.0, 10.0, 10.0, z["ranges"], scan.pose_table(z["poses"]),
                             scan.beam_table(-math.pi, math.pi, N), 30.0)
    score, pmap = corc.grid_finalize(hit, miss)
    np.testing.assert_allclose(score, z["datamap"], rtol=1e-12, atol=0)
    assert np.array_equal(pmap, z["pmap"])


# ----------------------------------------------------------------------------- next rows: f-2 pose chain, f-4 virtual scan

def test_pose_chain_oracle_matches_reference_publishResult():
    """pyref.compose_pose and the host form scan.compose_odometry against sensor_sta after every call of the
    reference's own ICP.publishResult ([ICP]:181-190) on a 400-transform stream (golden)."""
    import b2slam.scan as scan
    z = load_golden("next_rows.npz")
    T, want = z["chain_T"], z["chain_traj"]
    st = tuple(z["chain_start"])
    assert st == (0.25, -1.5, 0.4)
    for k in range(T.shape[0]):
        st = pyref.compose_pose(st, T[k])
        assert st == tuple(want[k + 1]), k                    # the literal port is bit-identical, step by step
    traj = scan.compose_odometry(tuple(z["chain_start"]), T)
    assert np.array_equal(traj, want)


def test_virtual_scan_oracle_matches_reference_laserEstimation():
    z = load_golden("next_rows.npz")
    for i in range(int(z["vscan_count"])):
        g = lambda k: z["vscan%d_%s" % (i, k)]
        got = pyref.virtual_scan([g("obs_x"), g("obs_y")], g("pose"), float(g("angle_min")), float(g("angle_increment")),
                                 int(g("beams")))
        assert np.array_equal(got, g("ranges")), i
