"""GPU: the drop-in modules (dropin/icp.py, mapping.py, bresenham.py) imported under the reference's module names
and driven the way slam_ekf.py's laserCallback drives them ([SLAM]:63-95), against the oracle scan by scan."""
import importlib
import math
import os
import sys

import numpy as np
import pytest

from conftest import ROOT

pytestmark = pytest.mark.gpu

torch = pytest.importorskip("torch")


def _import_like_the_node():
    drop = os.path.join(ROOT, "dropin")
    saved = {k: sys.modules.pop(k, None) for k in ("icp", "mapping", "bresenham")}
    sys.path.insert(0, drop)
    try:
        mods = [importlib.import_module(k) for k in ("icp", "mapping", "bresenham")]   # `from icp import ICP` ...
    finally:
        sys.path.remove(drop)
        for k, v in saved.items():
            sys.modules.pop(k, None)
            if v is not None:
                sys.modules[k] = v
    return mods


def test_node_flow_through_the_dropin_modules():
    from b2slam import _lib, scan, synth
    from oracle import pyref
    if _lib.device_count() <= 0:
        pytest.fail("no CUDA device: the gpu tests must run on the B200 box")
    icp_mod, mapping_mod, drawing = _import_like_the_node()
    icp = icp_mod.ICP()                                  # [SLAM]:35  no arguments, parameters as the node reads them
    mapping = mapping_mod.Mapping(200, 200, 0.1)         # [SLAM]:33 with mapping.launch:14-16
    assert (icp.max_iter, icp.tolerance) == (30, 0.001) and mapping.pmap.shape == (200, 200)

    xy, _ = synth.room_sequence(77, 9, 120)              # the simulator's 120 beams ([GAZ]:39)
    ranges = np.hypot(xy[:, 0], xy[:, 1]).astype(np.float32)
    ranges[3, 7] = np.inf                                # a beam without return: clamped to 30 m ([SLAM]:119)
    datamap = np.zeros((200, 200))
    pmap_ref = np.full((200, 200), 50.0)
    state = (0.5, -1.0, 0.3)
    tar_pc = None
    for k in range(ranges.shape[0]):
        np_msg = scan.laser_to_points(ranges[k], -math.pi, math.pi, clamp_inf_to=30)          # laserToNumpy
        if tar_pc is None:                                                                    # [SLAM]:74-78
            tar_pc = np_msg
            continue
        T = icp.process(tar_pc, np_msg)                                                       # calc_odometry
        want_T, want_it = pyref.icp_process(tar_pc, np_msg, 30, 1e-3)
        np.testing.assert_allclose(T, want_T, rtol=0, atol=1e-9)
        assert icp.last_iterations == want_it
        tar_pc = np_msg
        state = pyref.compose_pose(state, T)                                                  # odometry only (no EKF here)
        obs = scan.u2T(np.array(state)).dot(np_msg)                                           # [SLAM]:89
        pmap = mapping.update(obs[0], obs[1], np.array([state[0]]), np.array([state[1]]))     # [SLAM]:90 (1-element arrays)
        # float64 straight through, as the reference consumes it ([MAP]:33-36): nothing is rounded on either side
        pyref.grid_update_evidence(datamap, pmap_ref, obs[0], obs[1], state[0], state[1], 10.0, 10.0, 10.0)
        assert np.array_equal(pmap, pmap_ref), k
        data = np.trunc(np.asarray(list(pmap.T.reshape(-1)))).astype(np.int8)                 # publishMap, [SLAM]:270-271
        assert data.shape == (40000,) and set(np.unique(data)) <= {0, 50, 100}
    np.testing.assert_allclose(mapping.datamap, datamap, rtol=1e-5, atol=0)
    assert drawing.bresenham([3, 4], [17, -9]).path == [tuple(c) for c in pyref.bresenham_cells([3, 4], [17, -9])]
