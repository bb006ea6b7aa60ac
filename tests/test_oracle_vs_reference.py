"""Live pinning: run the UNMODIFIED reference classes (shim-loaded from /root/reference) beside the
oracle on fresh seeded inputs.  Skipped where the reference tree does not exist (the GPU box); the
committed golden vectors (tests/golden/, test_oracle.py) cover that case."""
import numpy as np
import pytest

from oracle import corc, pyref, ref_loader
import b2slam.synth as synth

pytestmark = pytest.mark.skipif(not ref_loader.available(), reason="reference tree not mounted")


def test_reference_bresenham_equals_oracle_on_random_segments():
    _, ref_bres = ref_loader.load_mapping_classes()
    rng = np.random.Generator(np.random.PCG64(424242))
    for _ in range(300):
        a = [int(v) for v in rng.integers(-200, 900, size=2)]
        b = [int(v) for v in rng.integers(-200, 900, size=2)]
        want = ref_bres(list(a), list(b)).path
        assert corc.bresenham(a, b) == want
        assert pyref.bresenham_cells(a, b) == want


@pytest.mark.parametrize("loader,w_hit", [(ref_loader.load_mapping_classes, 20.0),
                                           (ref_loader.load_mapping_online_classes, 4.0)])
def test_reference_mapping_equals_oracle_counts(loader, w_hit):
    Mapping, _ = loader()
    ref = Mapping(200, 200, 0.1)
    hit = np.zeros((200, 200), dtype=np.int32)
    miss = np.zeros((200, 200), dtype=np.int32)
    ox, oy, cx, cy = synth.grid_scans(31337, 6, 200, half_extent_m=9.0)
    for k in range(6):
        pm = ref.update(ox[k].astype(np.float64), oy[k].astype(np.float64), float(cx[k]), float(cy[k]))
        corc.grid_raycast(hit, miss, 10.0, 10.0, 10.0, ox[k][None], oy[k][None], cx[k:k + 1], cy[k:k + 1])
    score, pmap = corc.grid_finalize(hit, miss, w_hit)
    np.testing.assert_allclose(score, ref.datamap, rtol=1e-12, atol=0)
    ok = (pmap == np.asarray(pm).astype(np.int8)) | pyref.boundary_ambiguous(hit, miss, w_hit)
    assert ok.all()
    if w_hit == 20.0:
        assert np.array_equal(pmap, np.asarray(pm).astype(np.int8))


def test_reference_mapping_on_unrounded_float64_equals_oracle():
    """Fresh float64 coordinates (never rounded to float32), incl. decimal cell boundaries, through the unmodified
    reference and the float64 oracle entry."""
    Mapping, _ = ref_loader.load_mapping_classes()
    ref = Mapping(200, 200, 0.1)
    rng = np.random.Generator(np.random.PCG64(20260001))
    hit = np.zeros((200, 200), dtype=np.int32)
    miss = np.zeros((200, 200), dtype=np.int32)
    for k in range(5):
        cx, cy = rng.uniform(-9, 9, size=2)
        ox = np.concatenate([np.round(rng.uniform(-9.9, 9.9, 60), 1), rng.uniform(-12, 12, 60)])
        oy = np.concatenate([np.round(rng.uniform(-9.9, 9.9, 60), 1), rng.uniform(-12, 12, 60)])
        pm = ref.update(ox, oy, cx, cy)
        corc.grid_raycast(hit, miss, 10.0, 10.0, 10.0, ox[None], oy[None], [cx], [cy])
    score, pmap = corc.grid_finalize(hit, miss, 20.0)
    np.testing.assert_allclose(score, ref.datamap, rtol=1e-12, atol=0)
    assert np.array_equal((hit + miss) > 0, np.asarray(ref.datamap) > 0)
    assert np.array_equal(pmap, np.asarray(pm).astype(np.int8))


def test_reference_icp_equals_oracle_on_fresh_pairs():
    tar, src, _ = synth.icp_pairs(99001, 3, 120)
    for p in range(3):
        icp = ref_loader.load_icp_class({})()
        T = icp.process(synth.homogeneous(tar[p].astype(np.float64)), synth.homogeneous(src[p].astype(np.float64)))
        got, _ = corc.icp_batch(tar[p][None], src[p][None], 30, 1e-3)
        np.testing.assert_allclose(got[0], T, rtol=0, atol=1e-12)


def test_fhb_peer_copy_agrees_where_it_can_run():
    """icp-fhb.py is a pure-NumPy peer copy; it crashes in the reflection branch, so only well-posed pairs."""
    tar, src, _ = synth.icp_pairs(99002, 2, 120)
    fhb = ref_loader.load_fhb_icp_class()()
    for p in range(2):
        T = fhb.process(synth.homogeneous(tar[p].astype(np.float64)), synth.homogeneous(src[p].astype(np.float64)))
        got, _ = corc.icp_batch(tar[p][None], src[p][None], 30, 1e-3)
        np.testing.assert_allclose(got[0], T, rtol=0, atol=1e-12)
