"""CPU oracle, literal form: a float64 Python/NumPy restatement of the reference hot path.

TEST INFRASTRUCTURE ONLY -- never imported by the product package.  Allowed users:
tests/, __graft_entry__.smoke(), and bench.py's cpu_baseline / --impl reference legs.

It follows the reference's arithmetic operation by operation (same NumPy calls in
the same order), so it is both the parity checker for small cases and the honest
"what the reference costs on a CPU" baseline: the reference itself is pure
Python loops around np.linalg.norm, and so is this.

Citations are to /root/reference (aliases as in SURVEY.md):
  [ICP]  W9_Fusion Localization (LiDAR Odometry)/course_agv_slam/scripts/icp.py
  [MAP]  W12_LiDAR SLAM/w12-mapping/course_agv_slam/scripts/mapping.py
  [MAPO] W12_LiDAR SLAM/w12-mapping-online/course_agv_slam/scripts/mapping.py
  [BRES] W12_LiDAR SLAM/w12-mapping/course_agv_slam/scripts/bresenham.py

Pinning: tests/test_oracle_vs_reference.py executes the unmodified reference
classes (oracle/ref_loader.py) beside these functions whenever /root/reference is
present, and tests/golden/*.npz (written by oracle/make_golden.py from the same
reference classes) pin them where it is not.
"""
import math

import numpy as np


# --------------------------------------------------------------------------- ICP

def nearest_targets(src, tar):
    """Brute-force nearest neighbour, [ICP]:90-114.

    src (N,2), tar (M,2) float64 -> (distances (N,), indices (N,)).
    Strict `<` on the Euclidean norm, ascending j: the lowest index wins ties,
    NaN distances never win, an all-NaN row keeps index 0 / distance 0.
    """
    n = src.shape[0]
    which = np.zeros(n, dtype=np.int64)
    how_far = np.zeros(n)
    for i in range(n):
        p = src[i]
        best = np.inf
        for j in range(tar.shape[0]):
            d = np.linalg.norm(p - tar[j])
            if d < best:
                best = d
                which[i] = j
                how_far[i] = d
    return how_far, which


def rigid_fit_svd(src, tar):
    """Kabsch fit of row-matched point sets, [ICP]:149-179 (SVD form, W9 reflection fix).

    src (N,2), tar (N,2) -> 3x3 T with tar ~= R src + t.
    """
    ca = np.mean(src, axis=0)
    cb = np.mean(tar, axis=0)
    a0 = src - ca
    b0 = tar - cb
    w = np.dot(b0.transpose(), a0)
    u, _, vt = np.linalg.svd(w)
    r = np.dot(u, vt)
    if np.linalg.det(r) < 0:
        vt[1, :] *= -1
        r = np.dot(u, vt)
    t = cb.transpose() - np.dot(r, ca.transpose())
    out = np.identity(3)
    out[:2, :2] = r
    out[0, 2] = t[0]
    out[1, 2] = t[1]
    return out


def rigid_fit_closed_form(src, tar):
    """Same fit with the SVD replaced by its 2-D closed form.

    For W = sum bb_i aa_i^T the proper rotation U V^T (with the reflection fix of
    [ICP]:164-169) is the rotation by atan2(W10 - W01, W00 + W11)  (SURVEY.md section 7,
    verified against the SVD form in tests/test_oracle.py).  This is the form the
    CUDA kernel evaluates.  A vanishing W (hypot == 0) yields the identity rotation;
    the reference returns an arbitrary orthogonal matrix there.
    """
    ca = np.mean(src, axis=0)
    cb = np.mean(tar, axis=0)
    a0 = src - ca
    b0 = tar - cb
    w = np.dot(b0.transpose(), a0)
    cc = w[0, 0] + w[1, 1]
    ss = w[1, 0] - w[0, 1]
    h = math.hypot(cc, ss)
    if h > 0.0:
        c, s = cc / h, ss / h
    else:
        c, s = 1.0, 0.0
    out = np.identity(3)
    out[0, 0] = c
    out[0, 1] = -s
    out[1, 0] = s
    out[1, 1] = c
    out[0, 2] = cb[0] - (c * ca[0] - s * ca[1])
    out[1, 2] = cb[1] - (s * ca[0] + c * ca[1])
    return out


def icp_process(tar_pc, src_pc, max_iter=30, tolerance=1e-3, fit=rigid_fit_svd,
                nearest=nearest_targets):
    """ICP.process, [ICP]:38-88.  tar_pc (3,M) / src_pc (3,N) homogeneous float64.

    Returns (T 3x3, iterations run).  The error compared against `tolerance` is the
    mean NN distance measured BEFORE the iteration's move ([ICP]:75); the strict `<`
    means tolerance=0 never breaks; the returned T is a final re-fit of the original
    source onto the moved source ([ICP]:81).
    """
    first = np.array(src_pc[:2, :], dtype=np.float64)
    tar = np.ones((3, tar_pc.shape[1]))
    tar[:2, :] = tar_pc[:2, :]
    cur = np.ones((3, first.shape[1]))
    cur[:2, :] = first
    prev_err = 0
    done = 0
    for _ in range(max_iter):
        dist, idx = nearest(cur[:2, :].transpose(), tar[:2, :].transpose())
        step = fit(cur[:2, :].transpose(), tar[:2, idx].transpose())
        cur = np.dot(step, cur)
        done += 1
        err = np.sum(dist) / dist.size
        if abs(prev_err - err) < tolerance:
            break
        prev_err = err
    return fit(first.transpose(), cur[:2, :].transpose()), done


def nearest_targets_vec(src, tar):
    """Vectorised nearest_targets (same float64 values, first-index argmin)."""
    dx = src[:, None, 0] - tar[None, :, 0]
    dy = src[:, None, 1] - tar[None, :, 1]
    d = np.sqrt(dx * dx + dy * dy)
    which = np.argmin(d, axis=1)
    return d[np.arange(src.shape[0]), which], which.astype(np.int64)


# ---------------------------------------------------------------- Bresenham / grid

def bresenham_cells(start, end):
    """bresenham(start, end).path, [BRES]:2-58, as a list of (x, y) tuples.

    The trace direction is canonical (ascending major axis after the steep swap);
    `error` is a float64 accumulator of dy/float(dx) stepped at >= 0.5 -- NOT the
    integer algorithm (they differ on ~16% of slopes).
    """
    ax, ay = int(start[0]), int(start[1])
    bx, by = int(end[0]), int(end[1])
    if ax == bx and ay == by:
        return []
    steep = abs(by - ay) > abs(bx - ax)
    if steep:
        ax, ay = ay, ax
        bx, by = by, bx
    flipped = ax > bx
    if flipped:
        ax, bx = bx, ax
        ay, by = by, ay
    span = bx - ax
    rise = abs(by - ay)
    slope = rise / float(span)
    acc = 0.0
    minor = ay
    inc = 1 if ay < by else -1
    cells = []
    for major in range(ax, bx + 1):
        cells.append((minor, major) if steep else (major, minor))
        acc += slope
        if acc >= 0.5:
            minor += inc
            acc -= 1.0
    if flipped:
        cells.reverse()
    return cells


def world_to_cell(v, cells_per_m, offset_m):
    """int(S * (v + H)) in float64 with truncation toward zero, [MAP]:33-36.

    The reference hard-codes S = 10, H = 10 (its 200x200 / 0.1 m map); the general
    form S = 1/xyreso, H = extent/2 evaluates to exactly those literals there.
    """
    return int(cells_per_m * (float(v) + offset_m))


def grid_scale(xw, yw, xyreso):
    """(S, Hx, Hy) for a Mapping(xw, yw, xyreso); (10.0, 10.0, 10.0) at (200, 200, 0.1)."""
    return 1.0 / xyreso, xw * xyreso / 2.0, yw * xyreso / 2.0


def grid_update_counts(hit, miss, ox, oy, cx, cy, cells_per_m, off_x, off_y):
    """Integer form of Mapping.update, [MAP]:22-51: per-cell endpoint hits / traversals.

    hit, miss: int32 (xw, yw) arrays indexed [x][y], updated in place.
    ox, oy: (N,) endpoints; cx, cy: sensor position.  Beams with infinite ox are
    skipped ([MAP]:30); each path cell inside the grid counts one `miss` unless it
    is the LAST path element (the endpoint), which counts one `hit`; out-of-grid
    cells are skipped one by one; a same-cell beam touches nothing.
    Returns the number of in-grid cell visits (the V of SURVEY.md section 8d).
    """
    xw, yw = hit.shape
    visits = 0
    pcx = world_to_cell(cx, cells_per_m, off_x)
    pcy = world_to_cell(cy, cells_per_m, off_y)
    for i in range(len(ox)):
        if np.isinf(ox[i]):
            continue
        pox = world_to_cell(ox[i], cells_per_m, off_x)
        poy = world_to_cell(oy[i], cells_per_m, off_y)
        cells = bresenham_cells([pcx, pcy], [pox, poy])
        last = len(cells) - 1
        for j, (px, py) in enumerate(cells):
            if 0 <= px < xw and 0 <= py < yw:
                if j < last:
                    miss[px, py] += 1
                else:
                    hit[px, py] += 1
                visits += 1
    return visits


def grid_update_evidence(datamap, pmap, ox, oy, cx, cy, cells_per_m, off_x, off_y,
                         w_hit=20.0, w_miss=0.01, thresh=10.0):
    """Floating form of Mapping.update exactly as the reference runs it ([MAP]:39-50).

    Sequential float64 accumulation of +w_miss / +w_hit and the per-visit threshold;
    w_hit = 20 is [MAP]:45, w_hit = 4 is [MAPO]:46.
    """
    xw, yw = datamap.shape
    pcx = world_to_cell(cx, cells_per_m, off_x)
    pcy = world_to_cell(cy, cells_per_m, off_y)
    for i in range(len(ox)):
        if np.isinf(ox[i]):
            continue
        pox = world_to_cell(ox[i], cells_per_m, off_x)
        poy = world_to_cell(oy[i], cells_per_m, off_y)
        cells = bresenham_cells([pcx, pcy], [pox, poy])
        last = len(cells) - 1
        for j, (px, py) in enumerate(cells):
            if 0 <= px < xw and 0 <= py < yw:
                datamap[px, py] += w_miss if j < last else w_hit
                pmap[px, py] = 100 if datamap[px, py] > thresh else 0
    return pmap


def finalize_counts(hit, miss, w_hit=20.0, w_miss=0.01, thresh=10.0):
    """counts -> (datamap float64, pmap int8 in {0, 50, 100}), SURVEY.md section 8a row A6.

    datamap = w_miss*m + w_hit*h (product form); pmap = 50 where untouched, else
    100 if datamap > thresh else 0.  Equal to the reference's sequential sum for
    w_hit = 20 always; for w_hit = 4 equal except at exact-threshold counts where the
    reference itself depends on visit order (boundary_ambiguous()).
    """
    h = hit.astype(np.float64)
    m = miss.astype(np.float64)
    score = w_miss * m + w_hit * h
    pm = np.where(score > thresh, 100, 0).astype(np.int8)
    pm[(hit == 0) & (miss == 0)] = 50
    return score, pm


def boundary_ambiguous(hit, miss, w_hit=20.0, w_miss=0.01, thresh=10.0, rel=1e-9):
    """Cells whose score sits on the threshold to within accumulated rounding."""
    score = w_miss * miss.astype(np.float64) + w_hit * hit.astype(np.float64)
    return np.abs(score - thresh) <= rel * max(abs(thresh), 1.0)


# ------------------------------------------------------------------ adjacent steps

def laser_to_points(ranges, angle_min, angle_max, clamp_inf_to=None):
    """laserToNumpy, [ICP]:216-229 (no clamp) / W12 slam_ekf.py:115-123 (inf -> 30 m)."""
    r = np.array(ranges, dtype=np.float64)
    if clamp_inf_to is not None:
        r[r == np.inf] = clamp_inf_to
    n = r.shape[0]
    ang = np.linspace(angle_min, angle_max, n)
    pc = np.ones((3, n))
    pc[0, :] = np.cos(ang) * r
    pc[1, :] = np.sin(ang) * r
    return pc


def scan_to_world(ranges, pose, angle_min, angle_max, clamp_inf_to=30):
    """W12 slam_ekf.py:89: obs = u2T(xEst[:3]).dot(laserToNumpy(msg)) -> world-frame (ox, oy), float64."""
    x, y, w = (float(v) for v in pose)
    pc = laser_to_points(ranges, angle_min, angle_max, clamp_inf_to)
    t2 = np.array([[math.cos(w), -math.sin(w), x], [math.sin(w), math.cos(w), y]])  # u2T, slam_ekf.py:130-137
    obs = t2.dot(pc)
    return obs[0], obs[1]


def compose_pose(state, t_mat):
    """One step of the odometry chain, [ICP]:185-190 / W9 localization.py:79-83."""
    x, y, th = state
    dyaw = math.atan2(t_mat[1, 0], t_mat[0, 0])
    nx = x + math.cos(th) * t_mat[0, 2] - math.sin(th) * t_mat[1, 2]
    ny = y + math.sin(th) * t_mat[0, 2] + math.cos(th) * t_mat[1, 2]
    return (nx, ny, th + dyaw)


def virtual_scan(obstacle_xy, pose, angle_min, angle_increment, beams, far_range=100.0):
    """laserEstimation, W9 localization.py:128-150: min range per bearing bin over the obstacle cells."""
    x, y, yaw = (float(v) for v in pose)
    ranges = [far_range] * beams
    for i in range(len(obstacle_xy[0])):
        dist = math.hypot(x - obstacle_xy[0][i], y - obstacle_xy[1][i])
        index = int((math.atan2(obstacle_xy[1][i] - y, obstacle_xy[0][i] - x) - angle_min - yaw) / angle_increment)
        while index > beams - 1:
            index = index - beams
        while index < 0:
            index = index + beams
        if dist < ranges[index]:
            ranges[index] = dist
    return np.array(ranges)
