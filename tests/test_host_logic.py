"""Host-side pieces (no GPU): synthetic workloads, scan conversion, pose chain, sharding."""
import math

import numpy as np

from oracle import pyref
import b2slam.synth as synth
import b2slam.scan as scan
import b2slam.dist as bdist


def test_synth_is_seeded_and_float32():
    a = synth.icp_pairs(4001, 3, 360)
    b = synth.icp_pairs(4001, 3, 360)
    assert all(np.array_equal(x, y) for x, y in zip(a, b))
    assert a[0].dtype == np.float32 and a[0].shape == (3, 2, 360)
    ox, oy, cx, cy = synth.grid_scans(12001, 16, 1080)
    assert ox.shape == (16, 1080) and ox.dtype == np.float32 and cx.shape == (16,)
    r = np.hypot(ox - cx[:, None], oy - cy[:, None])
    assert r.min() >= 0.09 and r.max() <= 30.01
    assert np.abs(cx).max() <= 80.0 and np.abs(cy).max() <= 80.0


def test_room_sequence_steps_are_small():
    xy, poses = synth.room_sequence(9001, 40, 360)
    assert xy.shape == (40, 2, 360) and np.isfinite(xy).all()
    d = np.hypot(np.diff(poses[:, 0]), np.diff(poses[:, 1]))
    assert d.max() <= 0.1 + 1e-12
    assert np.abs(np.diff(poses[:, 2])).max() <= 0.05 + 1e-12
    rng = np.hypot(xy[:, 0], xy[:, 1])
    assert rng.max() <= 30.0 and rng.min() >= 0.1 - 1e-6


def test_laser_to_points_matches_oracle_and_clamps():
    rng = np.random.Generator(np.random.PCG64(3))
    ranges = rng.uniform(0.1, 30, 120).astype(np.float32)
    ranges[5] = np.inf
    a = scan.laser_to_points(ranges, -math.pi, math.pi, clamp_inf_to=scan.MAX_LASER_RANGE)
    b = pyref.laser_to_points(ranges, -math.pi, math.pi, clamp_inf_to=30)
    assert np.array_equal(a, b) and a.shape == (3, 120)
    assert abs(np.hypot(a[0, 5], a[1, 5]) - 30.0) < 1e-12
    assert np.isinf(scan.laser_to_points(ranges, -math.pi, math.pi)[:2, 5]).any()


def test_pose_chain_matches_oracle():
    rng = np.random.Generator(np.random.PCG64(4))
    Ts = []
    for _ in range(50):
        th = rng.uniform(-0.05, 0.05)
        T = np.identity(3)
        T[:2, :2] = [[math.cos(th), -math.sin(th)], [math.sin(th), math.cos(th)]]
        T[:2, 2] = rng.uniform(-0.1, 0.1, 2)
        Ts.append(T)
    traj = scan.compose_odometry((0.0, 0.0, 0.3), Ts)
    st = (0.0, 0.0, 0.3)
    for i, T in enumerate(Ts):
        st = pyref.compose_pose(st, T)
        assert np.allclose(traj[i + 1], st, rtol=0, atol=1e-15)


def test_u2T_T2u_roundtrip():
    u = np.array([[1.5], [-2.0], [0.7]])
    T = np.identity(3)
    T[:2, :] = scan.u2T(u)
    assert np.allclose(scan.T2u(T), u)


def test_shard_bounds_cover_everything_once():
    for total in (0, 1, 7, 9999, 10 ** 6):
        for world in (1, 2, 3, 4, 8):
            spans = [bdist.shard_bounds(total, r, world) for r in range(world)]
            assert spans[0][0] == 0 and spans[-1][1] == total
            assert all(spans[i][1] == spans[i + 1][0] for i in range(world - 1))
            sizes = [b - a for a, b in spans]
            assert max(sizes) - min(sizes) <= 1
    assert bdist.sequence_pair_bounds(10000, 0, 1) == (0, 9999)


def test_pose_table_uses_the_reference_trigonometry():
    """u2T calls math.cos / math.sin; the vectorised table must hold the very same doubles."""
    rng = np.random.Generator(np.random.PCG64(8))
    poses = np.stack([rng.uniform(-80, 80, 20000), rng.uniform(-80, 80, 20000), rng.uniform(-50, 50, 20000)], axis=1)
    tab = scan.pose_table(poses)
    assert np.array_equal(tab[:, 2], np.array([math.cos(w) for w in poses[:, 2]]))
    assert np.array_equal(tab[:, 3], np.array([math.sin(w) for w in poses[:, 2]]))
    assert np.array_equal(tab[:, :2], poses[:, :2])
    cs = scan.beam_table(-math.pi, math.pi, 1080)
    a = np.linspace(-math.pi, math.pi, 1080)
    assert np.array_equal(cs[:, 0], np.cos(a)) and np.array_equal(cs[:, 1], np.sin(a))
