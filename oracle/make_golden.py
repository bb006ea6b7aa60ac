"""Generate tests/golden/*.npz by EXECUTING the unmodified reference classes.

Run in the build container (where /root/reference is mounted):
    python oracle/make_golden.py
The reference has no tests, fixtures or golden vectors of its own, so
these files -- reference inputs and the reference's own outputs -- are what pins the oracle
and, through it, the CUDA path on machines where /root/reference does not exist.

TEST INFRASTRUCTURE ONLY.
"""
import importlib.util
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
sys.path.insert(0, ROOT)

from oracle import ref_loader  # noqa: E402

_spec = importlib.util.spec_from_file_location(
    "_b2s_synth",
    os.path.join(ROOT, "a-2d-lidar-based-slam-system-for-wheeled-mobile-robots_b200", "synth.py"))
synth = importlib.util.module_from_spec(_spec)
_spec.loader.exec_module(synth)

OUT = os.path.join(ROOT, "tests", "golden")


def _counting_icp(params):
    cls = ref_loader.load_icp_class(params)
    obj = cls()
    calls = {"n": 0}
    inner = obj.findNearest

    def counted(src, tar):
        calls["n"] += 1
        return inner(src, tar)

    obj.findNearest = counted
    return obj, calls


def icp_pairs_golden():
    """(i) whole-pipeline ICP.process on seeded pairs, both parameter sets."""
    cases = []
    # (seed, beams, pairs, max_iter, tolerance)
    plan = [
        (7101, 120, 6, 30, 0.001),   # real simulator beam count, node defaults
        (7102, 120, 3, 10, 0.0),     # W7 launch parameters: never breaks early
        (7001, 360, 1, 10, 0.0),     # cfg 1 as launched by W7 icp.launch
        (7001, 360, 1, 30, 0.001),   # cfg 1 with the defaults W8/W9/W12 silently use
        (7103, 360, 1, 30, 0.001),
    ]
    for seed, beams, pairs, max_iter, tol in plan:
        tar, src, truth = synth.icp_pairs(seed, pairs, beams)
        for p in range(pairs):
            icp, calls = _counting_icp({"/icp/tolerance": tol, "/icp/max_iter": max_iter})
            assert icp.max_iter == max_iter
            T = icp.process(synth.homogeneous(tar[p].astype(np.float64)),
                            synth.homogeneous(src[p].astype(np.float64)))
            cases.append(dict(tar=tar[p], src=src[p], T=np.asarray(T, dtype=np.float64),
                              iters=calls["n"], max_iter=max_iter, tol=tol, seed=seed,
                              truth=truth[p]))
            print("icp seed=%d beams=%d pair=%d iters=%d" % (seed, beams, p, calls["n"]))
    # unequal cloud sizes (N != M): the virtual-scan ICP of W9 localization.py:152-157
    tar, src, _ = synth.icp_pairs(7104, 1, 150)
    icp, calls = _counting_icp({})
    T = icp.process(synth.homogeneous(tar[0].astype(np.float64)),
                    synth.homogeneous(src[0][:, :97].astype(np.float64)))
    cases.append(dict(tar=tar[0], src=src[0][:, :97], T=np.asarray(T), iters=calls["n"],
                      max_iter=30, tol=0.001, seed=7104, truth=np.zeros(3)))
    blob = {"count": len(cases)}
    for i, c in enumerate(cases):
        for k, v in c.items():
            blob["%d_%s" % (i, k)] = v
    np.savez_compressed(os.path.join(OUT, "icp_pairs.npz"), **blob)


def _reference_pair(args):
    """One whole ICP.process of the unmodified reference (worker of icp_pairs_1080_golden)."""
    tar, src, max_iter, tol = args
    icp, calls = _counting_icp({"/icp/tolerance": tol, "/icp/max_iter": max_iter})
    T = icp.process(synth.homogeneous(tar.astype(np.float64)), synth.homogeneous(src.astype(np.float64)))
    return np.asarray(T, dtype=np.float64), calls["n"]


def icp_pairs_1080_golden(pairs=6):
    """(i, cfg 4 shape) the first pairs of the cfg-4 stream (seed 4001, 1080 beams) through the reference itself:
    ~1.2 M np.linalg.norm calls per iteration, about half a minute per pair, so the pairs run in parallel."""
    import multiprocessing as mp
    tar, src, truth = synth.icp_pairs(4001, pairs, 1080)
    with mp.Pool(min(pairs, os.cpu_count() or 1)) as pool:
        res = pool.map(_reference_pair, [(tar[p], src[p], 30, 0.001) for p in range(pairs)])
    blob = {"count": pairs}
    for i, (T, iters) in enumerate(res):
        for k, v in dict(tar=tar[i], src=src[i], T=T, iters=iters, max_iter=30, tol=0.001, seed=4001,
                         truth=truth[i]).items():
            blob["%d_%s" % (i, k)] = v
        print("icp seed=4001 beams=1080 pair=%d iters=%d" % (i, iters))
    np.savez_compressed(os.path.join(OUT, "icp_pairs_1080.npz"), **blob)


def nearest_and_fit_golden():
    """(ii) findNearest incl. exact ties, (iii) getTransform incl. reflection-branch inputs."""
    rng = np.random.Generator(np.random.PCG64(8101))
    icp = ref_loader.load_icp_class({})()
    blob = {}
    # ties: duplicated targets and symmetric layouts (lowest index must win)
    tar = rng.uniform(-5, 5, size=(40, 2)).astype(np.float32).astype(np.float64)
    tar[17] = tar[3]
    tar[30] = tar[3]
    tar[25] = tar[9]
    src = rng.uniform(-5, 5, size=(60, 2)).astype(np.float32).astype(np.float64)
    src[0] = tar[3]                       # zero distance onto a triplicated target
    src[1] = (tar[9] + np.array([0.25, 0.0]))
    sym_t = np.array([[1.0, 0.0], [-1.0, 0.0], [0.0, 1.0], [0.0, -1.0], [1.0, 0.0]])
    sym_s = np.array([[0.0, 0.0], [0.5, 0.5], [-0.5, 0.5], [0.0, 2.0]])
    d, i = icp.findNearest(src, tar)
    blob.update(tie_src=src, tie_tar=tar, tie_dist=d, tie_idx=np.asarray(i, dtype=np.int64))
    d, i = icp.findNearest(sym_s, sym_t)
    blob.update(sym_src=sym_s, sym_tar=sym_t, sym_dist=d, sym_idx=np.asarray(i, dtype=np.int64))
    # fits: random rigid motions, pure reflections of the cloud (forces det(U Vt) < 0), noise
    fs, ft, fT = [], [], []
    for k in range(24):
        n = 50
        a = rng.normal(0, 2.0, size=(n, 2))
        th = rng.uniform(-np.pi, np.pi)
        c, s = np.cos(th), np.sin(th)
        b = a @ np.array([[c, s], [-s, c]]) + rng.uniform(-1, 1, size=2)
        if k % 3 == 1:
            b = b * np.array([1.0, -1.0])          # mirrored target -> reflection branch
        if k % 3 == 2:
            b = b + rng.normal(0, 0.3, size=b.shape)
        fs.append(a)
        ft.append(b)
        fT.append(np.asarray(icp.getTransform(a, b)))
    blob.update(fit_src=np.array(fs), fit_tar=np.array(ft), fit_T=np.array(fT))
    np.savez_compressed(os.path.join(OUT, "icp_pieces.npz"), **blob)


def nearest_ties_golden():
    """(ii-b) ties that only the reference's sqrt creates.  findNearest compares np.linalg.norm values ([ICP]:102-103),
    and sqrt maps neighbouring doubles onto one: src (0,0) with targets (1, 2^-26), (-1, 0) has squared distances
    1 + 2^-52 and 1 but BOTH norms are exactly 1.0, so the reference keeps index 0 where an argmin over squared
    distances returns 1.  10^4 generated cases of that kind + whole ICP.process runs whose
    first-iteration correspondences hinge on such ties."""
    rng = np.random.Generator(np.random.PCG64(8601))
    icp = ref_loader.load_icp_class({})()
    blob = {}
    d, i = icp.findNearest(np.array([[0.0, 0.0]]), np.array([[1.0, 2.0 ** -26], [-1.0, 0.0]]))
    blob.update(repro_src=np.array([[0.0, 0.0]]), repro_tar=np.array([[1.0, 2.0 ** -26], [-1.0, 0.0]]),
                repro_dist=d, repro_idx=np.asarray(i, dtype=np.int64))
    assert int(i[0]) == 0 and d[0] == 1.0
    # generated cases: one source point each, 6 targets: a near-tied pair (x, k*2^-26*x) / mirror image in
    # either index order, k = 0..3 (k = 0 exact tie, 1 sqrt-tie, 2-3 no tie), and four farther decoys
    n_cases = 10000
    srcs = np.zeros((n_cases, 2))
    tars = np.zeros((n_cases, 6, 2))
    for c in range(n_cases):
        x = float(np.float32(rng.uniform(0.5, 8.0)))
        k = int(rng.integers(0, 4))
        th = float(rng.uniform(-np.pi, np.pi)) if c % 2 else 0.0       # half axis-aligned, half rotated (inexact products)
        cs, sn = np.cos(th), np.sin(th)
        a = np.array([x, k * 2.0 ** -26 * x])
        b = np.array([-x, 0.0])
        R = np.array([[cs, -sn], [sn, cs]])
        a, b = R @ a, R @ b
        o = rng.uniform(-3, 3, size=2) if c % 3 == 0 else np.zeros(2)
        pts = [a + o, b + o] if rng.integers(0, 2) else [b + o, a + o]
        decoys = [o + (R @ np.array([x * f, x * g])) for f, g in ((1.5, 0.1), (-1.4, 0.3), (0.2, 1.7), (0.1, -1.3))]
        slot = rng.permutation(6)
        allp = pts + decoys
        # keep the relative order of the tied pair as drawn, shuffle where the pair sits among the decoys
        order = sorted(range(6), key=lambda q: slot[q] if q >= 2 else min(slot[0], slot[1]) + 0.1 * q)
        tars[c] = np.array([allp[q] for q in order])
        srcs[c] = o
    idx = np.zeros(n_cases, dtype=np.int64)
    dist = np.zeros(n_cases)
    for c in range(n_cases):
        dd, ii = icp.findNearest(srcs[c:c + 1], tars[c])
        idx[c], dist[c] = int(ii[0]), dd[0]
    # how many of them an argmin over the squared distance gets wrong: (a) the radicand NumPy's norm actually takes the
    # root of, fma(dy, dy, dx*dx) (exact rational arithmetic, rounded once); (b) the textbook dx*dx + dy*dy
    from fractions import Fraction
    diff = srcs[:, None, :] - tars
    q_fma = np.array([[float(Fraction(d[1]) * Fraction(d[1]) + Fraction(d[0] * d[0])) for d in row] for row in diff])
    q_sum = diff[..., 0] * diff[..., 0] + diff[..., 1] * diff[..., 1]
    assert np.array_equal(np.sqrt(q_fma)[np.arange(n_cases), idx], dist)   # the norm IS sqrt(fma(dy, dy, dx*dx)) here
    blob.update(gen_src=srcs, gen_tar=tars, gen_idx=idx, gen_dist=dist,
                gen_d2_argmin_differs=int((q_fma.argmin(1) != idx).sum()),
                gen_plain_sum_differs=int((np.sqrt(q_sum).argmin(1) != idx).sum()))
    print("nearest ties: of %d cases %d differ from an argmin over the radicand, %d from sqrt(dx*dx + dy*dy)" % (
        n_cases, blob["gen_d2_argmin_differs"], blob["gen_plain_sum_differs"]))
    # whole ICP.process (max_iter 1 and the defaults): source points ON the mirror axis of a target cloud whose mirror
    # images are displaced by sqrt-tie amounts, so the correspondences of iteration 1 are all decided by the tie rule
    cases = []
    for c in range(6):
        m = 24
        xs = np.array([float(np.float32(v)) for v in rng.uniform(0.5, 4.0, m)])
        ys = np.array([float(np.float32(v)) for v in np.linspace(-3, 3, m)])
        k = rng.integers(0, 3, size=m)
        right = np.stack([xs, ys + k * 2.0 ** -26 * xs], axis=1)
        left = np.stack([-xs, ys], axis=1)
        tar = np.empty((2 * m, 2))
        first_right = rng.integers(0, 2, size=m).astype(bool)
        tar[0::2] = np.where(first_right[:, None], right, left)
        tar[1::2] = np.where(first_right[:, None], left, right)
        src = np.stack([np.zeros(m), ys], axis=1)
        src = np.concatenate([src, rng.uniform(-1, 1, size=(5, 2)) + np.array([0.7, 0.0])])   # a few asymmetric points
        for max_iter, tol in ((1, 0.0), (30, 0.001)):
            obj, calls = _counting_icp({"/icp/tolerance": tol, "/icp/max_iter": max_iter})
            T = obj.process(synth.homogeneous(tar.T.copy()), synth.homogeneous(src.T.copy()))
            cases.append(dict(tar=tar.T.copy(), src=src.T.copy(), T=np.asarray(T, dtype=np.float64), iters=calls["n"],
                              max_iter=max_iter, tol=tol))
    blob["icp_count"] = len(cases)
    for n, cse in enumerate(cases):
        for key, v in cse.items():
            blob["icp%d_%s" % (n, key)] = v
    np.savez_compressed(os.path.join(OUT, "icp_ties.npz"), **blob)


def bresenham_golden():
    """(iv) rasteriser: every (dx,dy) with |dx|,|dy| <= 24 in all octants, a dx<=96 sweep of the
    first octant (where float64 != integer Bresenham shows up), long random segments."""
    _, bres = ref_loader.load_mapping_classes()
    segs = []
    for dx in range(-24, 25):
        for dy in range(-24, 25):
            segs.append((100, 100, 100 + dx, 100 + dy))
    for dx in range(25, 97):
        for dy in range(0, dx + 1):
            segs.append((7, -3, 7 + dx, -3 + dy))
    rng = np.random.Generator(np.random.PCG64(8201))
    for _ in range(400):
        x0, y0 = rng.integers(-50, 700, size=2)
        x1, y1 = rng.integers(-50, 700, size=2)
        segs.append((int(x0), int(y0), int(x1), int(y1)))
    segs = np.array(segs, dtype=np.int32)
    offs = [0]
    cells = []
    for x0, y0, x1, y1 in segs:
        path = bres([int(x0), int(y0)], [int(x1), int(y1)]).path
        cells.extend(path)
        offs.append(len(cells))
    np.savez_compressed(os.path.join(OUT, "bresenham.npz"), segs=segs,
                        offsets=np.array(offs, dtype=np.int64),
                        cells=np.array(cells, dtype=np.int32).reshape(-1, 2))
    print("bresenham segments=%d cells=%d" % (len(segs), len(cells)))


def _edge_beams(rng, n, cx, cy):
    """Endpoints that exercise every branch of Mapping.update around a sensor at (cx, cy)."""
    ang = np.linspace(-np.pi, np.pi, n)
    r = rng.uniform(0.3, 9.0, size=n)
    ox = cx + r * np.cos(ang)
    oy = cy + r * np.sin(ang)
    ox[0], oy[0] = cx + 0.004, cy + 0.003          # same cell as the sensor: no update at all
    ox[1], oy[1] = cx + 30.0, cy + 0.5             # endpoint far outside: ray clipped, no hit
    ox[2], oy[2] = -10.05, cy                      # x in (-10.1,-10): truncation toward zero -> cell 0
    ox[3], oy[3] = cx, -10.07
    ox[4], oy[4] = np.inf, cy                      # skipped ([MAP]:30)
    ox[5], oy[5] = -np.inf, cy                     # skipped as well (isinf)
    ox[6], oy[6] = 9.999, 9.999                    # last in-map cell
    ox[7], oy[7] = 10.0, 10.0                      # first out-of-map cell (index 200)
    ox[8], oy[8] = -25.0, -25.0                    # negative cells
    ox[9], oy[9] = cx + 0.1, cy                    # one-cell step
    return ox.astype(np.float32), oy.astype(np.float32)


def mapping_golden():
    """(v) Mapping.update on the reference's own 200x200 / 0.1 m map, (vi) +4 variant."""
    rng = np.random.Generator(np.random.PCG64(8301))
    blob = {}
    for tag, loader in (("w20", ref_loader.load_mapping_classes),
                        ("w4", ref_loader.load_mapping_online_classes)):
        Mapping, _ = loader()
        m = Mapping(200, 200, 0.1)
        centers = [(0.0, 0.0), (3.0, 3.0), (-9.96, 9.93), (2.7, 2.7), (-10.04, -10.02),
                   (0.05, 0.0), (0.05, 0.0), (0.05, 0.0)]
        oxs, oys = [], []
        for (cx, cy) in centers:
            ox, oy = _edge_beams(rng, 120, cx, cy)
            oxs.append(ox)
            oys.append(oy)
            pm = m.update(ox.astype(np.float64), oy.astype(np.float64),
                          float(np.float32(cx)), float(np.float32(cy)))
        blob[tag + "_ox"] = np.array(oxs)
        blob[tag + "_oy"] = np.array(oys)
        blob[tag + "_cx"] = np.array([c[0] for c in centers], dtype=np.float32)
        blob[tag + "_cy"] = np.array([c[1] for c in centers], dtype=np.float32)
        blob[tag + "_datamap"] = np.array(m.datamap)
        blob[tag + "_pmap"] = np.array(pm).astype(np.int8)
        print(tag, "occupied", int((np.array(pm) == 100).sum()), "free", int((np.array(pm) == 0).sum()))
    # (vi) threshold crossings of the +0.01 stream: 1000 traversals stay free, 1001 flip
    Mapping, _ = ref_loader.load_mapping_classes()
    m = Mapping(200, 200, 0.1)
    ox = np.array([1.05], dtype=np.float32)
    oy = np.array([0.05], dtype=np.float32)
    snaps = {}
    for k in range(1, 1003):
        pm = m.update(ox.astype(np.float64), oy.astype(np.float64), 0.05, 0.05)
        if k in (1000, 1001, 1002):
            snaps[k] = (float(m.datamap[105][100]), int(pm[105][100]))
    blob["miss_stream_counts"] = np.array(sorted(snaps), dtype=np.int64)
    blob["miss_stream_score"] = np.array([snaps[k][0] for k in sorted(snaps)])
    blob["miss_stream_pmap"] = np.array([snaps[k][1] for k in sorted(snaps)], dtype=np.int64)
    np.savez_compressed(os.path.join(OUT, "mapping.npz"), **blob)


def _sparse(datamap, pmap):
    """Non-zero cells of the reference's datamap (flat [x][y] indices), their scores and occupancies."""
    idx = np.flatnonzero(np.asarray(datamap).reshape(-1))
    return idx.astype(np.int64), np.asarray(datamap).reshape(-1)[idx], np.asarray(pmap).reshape(-1)[idx].astype(np.int8)


def mapping_f64_golden():
    """(v-f64) Mapping.update fed what slam_ekf.py:89-90 feeds it: UNROUNDED float64 endpoints and sensor
    positions, at three grid scales.  The reference hard-codes int(10*(v+10)) ([MAP]:33-36), so on the larger
    grids the world simply extends to 10*(v+10) < xw: coordinates up to ~400 m (4096^2) and ~1628 m (16384^2),
    where a float32 narrowing of the coordinate moves the cell of roughly one beam in 10^4 / 10^3."""
    blob = {}
    for tag, side, span, scans, beams, seed in (("g200", 200, 19.0, 8, 160, 8501), ("g4096", 4096, 405.0, 6, 200, 8502),
                                                ("g16384", 16384, 1630.0, 6, 200, 8503)):
        rng = np.random.Generator(np.random.PCG64(seed))
        Mapping, _ = ref_loader.load_mapping_classes()
        MappingO, _ = ref_loader.load_mapping_online_classes()
        m, mo = Mapping(side, side, 0.1), (MappingO(side, side, 0.1) if side == 200 else None)
        oxs, oys, cxs, cys = [], [], [], []
        for k in range(scans):
            # sensor somewhere in the (extended) world, not float32-representable
            cx = float(rng.uniform(-9.5, span - 10.5))
            cy = float(rng.uniform(-9.5, span - 10.5))
            if k == 1:
                cx, cy = 0.30000000000000004, 0.7                  # 0.1 + 0.2; 10*(0.7+10) = 106.99999999999999
            ang = np.linspace(-np.pi, np.pi, beams)
            r = rng.uniform(0.3, 28.0, size=beams)
            ox = cx + r * np.cos(ang)
            oy = cy + r * np.sin(ang)
            # endpoints exactly on the decimal cell boundaries: k/10 is not a binary fraction, and 10*(k/10+10)
            # falls on either side of the integer (9.9 -> 198.99999999999997 -> cell 198, not 199)
            nb = beams // 4
            ticks = np.round(np.clip(cx + rng.uniform(-25, 25, size=nb), -9.9, span - 10.1), 1)
            ox[:nb] = ticks
            oy[:nb] = np.round(cy + rng.uniform(-20, 20, size=nb), 1)
            # one ulp either side of a boundary
            ox[nb] = np.nextafter(np.round(cx + 3.0, 1), np.inf)
            ox[nb + 1] = np.nextafter(np.round(cx + 3.0, 1), -np.inf)
            ox[nb + 2], oy[nb + 2] = np.inf, cy                      # skipped ([MAP]:30)
            ox[nb + 3], oy[nb + 3] = cx + 1e-9, cy - 1e-9            # same cell: no update
            ox[nb + 4], oy[nb + 4] = -10.05, cy                      # truncation toward zero -> cell 0
            ox[nb + 5], oy[nb + 5] = cx, span + 50.0                 # leaves the grid: clipped, no hit
            pm = m.update(ox, oy, cx, cy)
            if mo is not None:
                pmo = mo.update(ox, oy, cx, cy)
            oxs.append(ox); oys.append(oy); cxs.append(cx); cys.append(cy)
        blob[tag + "_ox"], blob[tag + "_oy"] = np.array(oxs), np.array(oys)
        blob[tag + "_cx"], blob[tag + "_cy"] = np.array(cxs), np.array(cys)
        blob[tag + "_side"] = side
        blob[tag + "_cells"], blob[tag + "_score"], blob[tag + "_pmap"] = _sparse(m.datamap, pm)
        if mo is not None:
            blob[tag + "_cells_w4"], blob[tag + "_score_w4"], blob[tag + "_pmap_w4"] = _sparse(mo.datamap, pmo)
        # how many cells of these very scans move when the coordinates are narrowed to float32 first (informational)
        f = lambda v: np.trunc(10 * (np.asarray(v, dtype=np.float64) + 10))
        g = lambda v: np.trunc(10 * (np.asarray(v, dtype=np.float32).astype(np.float64) + 10))
        fin = np.isfinite(blob[tag + "_ox"])
        moved = int(((f(blob[tag + "_ox"][fin]) != g(blob[tag + "_ox"][fin])) | (f(blob[tag + "_oy"][fin]) != g(blob[tag + "_oy"][fin]))).sum())
        blob[tag + "_moved_by_f32"] = moved
        print("mapping f64 %s: %d touched cells, %d endpoints would move under float32 narrowing" % (
            tag, len(blob[tag + "_cells"]), moved))
        del m, mo
    np.savez_compressed(os.path.join(OUT, "mapping_f64.npz"), **blob)


class _Scan(object):
    """Duck-typed sensor_msgs/LaserScan."""

    def __init__(self, ranges, angle_min, angle_max):
        self.ranges = list(ranges)
        self.angle_min = angle_min
        self.angle_max = angle_max


def ingestion_golden():
    """(vii) the node's scan ingestion around Mapping.update: laserToNumpy (inf -> 30 m) + u2T(xEst).dot(np_msg)
    (slam_ekf.py:89-90, 115-137), executed with the reference's own functions on raw float32 ranges."""
    import math
    Node, Mapping = ref_loader.load_slam_node_class()
    rng = np.random.Generator(np.random.PCG64(8401))
    K, N = 8, 120
    ranges = np.clip(synth.noisy(rng, synth.clean_ranges(rng, K, N)), 0.1, 30.0).astype(np.float32)
    ranges[1, 7] = np.inf            # clamped to MAX_LASER_RANGE = 30: ray leaves the 20 m map
    ranges[3, 50:53] = np.inf
    poses = np.stack([rng.uniform(-6, 6, K), rng.uniform(-6, 6, K), rng.uniform(-math.pi, math.pi, K)], axis=1)
    m = Mapping(200, 200, 0.1)
    oxs, oys = [], []
    for k in range(K):
        np_msg = Node.laserToNumpy(None, _Scan(ranges[k].tolist(), -math.pi, math.pi))
        # the node passes xEst[:3] as a (3,1) column; NumPy >= 1.24 refuses to build u2T's 2x3 array from
        # 1-element arrays, so the pose goes in as three scalars -- same arithmetic
        xEst = poses[k]
        obs = Node.u2T(None, xEst[:3]).dot(np_msg)
        pm = m.update(obs[0], obs[1], float(xEst[0]), float(xEst[1]))
        oxs.append(obs[0])
        oys.append(obs[1])
    np.savez_compressed(os.path.join(OUT, "ingestion.npz"), ranges=ranges, poses=poses, angle_min=-math.pi,
                        angle_max=math.pi, ox=np.array(oxs), oy=np.array(oys), datamap=np.array(m.datamap),
                        pmap=np.array(pm).astype(np.int8))
    print("ingestion: occupied", int((np.array(pm) == 100).sum()))


def next_rows_golden():
    """(viii) the two "next" rows of SURVEY section 8f that had only the oracle behind them:
    f-2  the odometry chain of ICP.publishResult ([ICP]:181-190): the reference's own method, called T by T on a stream
         of transforms, sensor_sta read back after every call;
    f-4  the virtual scan Localization.laserEstimation (W9 localization.py:128-150): the reference's own function on a
         duck-typed node (obstacle cells, xEst) and scan message."""
    import math
    import types as _types
    rng = np.random.Generator(np.random.PCG64(8701))
    blob = {}
    # ---- f-2
    cls = ref_loader.load_icp_class({"/icp/robot_x": 0.25, "/icp/robot_y": -1.5, "/icp/robot_theta": 0.4})
    icp = cls()
    n = 400
    th = rng.uniform(-0.08, 0.08, n)
    th[::37] = rng.uniform(-math.pi, math.pi, len(th[::37]))       # a few large turns: the yaw leaves (-pi, pi] and keeps growing
    tx, ty = rng.uniform(-0.15, 0.15, n), rng.uniform(-0.15, 0.15, n)
    T = np.zeros((n, 3, 3))
    T[:, 0, 0], T[:, 0, 1], T[:, 0, 2] = np.cos(th), -np.sin(th), tx
    T[:, 1, 0], T[:, 1, 1], T[:, 1, 2] = np.sin(th), np.cos(th), ty
    T[:, 2, 2] = 1.0
    traj = [list(icp.sensor_sta)]
    for k in range(n):
        icp.publishResult(T[k])
        traj.append([float(v) for v in icp.sensor_sta])
    blob.update(chain_T=T, chain_start=np.array(traj[0], dtype=np.float64), chain_traj=np.array(traj, dtype=np.float64))
    # ---- f-4
    Localization = ref_loader.load_localization_class()
    cases = []
    for c, (beams, cells) in enumerate(((120, 900), (360, 4000), (1080, 20000), (90, 0))):
        amin = -math.pi if c != 1 else -2.0
        amax = math.pi if c != 1 else 2.5
        ox = rng.uniform(-12, 12, cells)
        oy = rng.uniform(-12, 12, cells)
        if cells:
            ox[:5], oy[:5] = 100.0, 100.0                       # farther than the 100.0 the empty bins hold
        pose = np.array([rng.uniform(-3, 3), rng.uniform(-3, 3), rng.uniform(-math.pi, math.pi)])
        msg = _types.SimpleNamespace(ranges=[1.0] * beams, angle_min=amin, angle_max=amax,
                                     angle_increment=(amax - amin) / (beams - 1))
        node = _types.SimpleNamespace(obstacle=[ox, oy], xEst=pose, target_laser=None)
        out = Localization.laserEstimation(node, msg, pose)
        cases.append(dict(obs_x=ox, obs_y=oy, pose=pose, angle_min=amin, angle_increment=msg.angle_increment, beams=beams,
                          ranges=np.array(out.ranges, dtype=np.float64)))
    blob["vscan_count"] = len(cases)
    for i, cse in enumerate(cases):
        for k, v in cse.items():
            blob["vscan%d_%s" % (i, k)] = v
    np.savez_compressed(os.path.join(OUT, "next_rows.npz"), **blob)
    print("next rows: chain of %d transforms, %d virtual scans" % (n, len(cases)))


def main():
    if not ref_loader.available():
        raise SystemExit("reference tree not found at %s" % ref_loader.REF_ROOT)
    os.makedirs(OUT, exist_ok=True)
    only = sys.argv[1:]
    if only:                      # python oracle/make_golden.py mapping_f64_golden ...
        for name in only:
            globals()[name]()
        return
    mapping_f64_golden()
    nearest_ties_golden()
    next_rows_golden()
    ingestion_golden()
    bresenham_golden()
    nearest_and_fit_golden()
    mapping_golden()
    icp_pairs_golden()
    icp_pairs_1080_golden()


if __name__ == "__main__":
    main()
